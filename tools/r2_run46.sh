#!/usr/bin/env bash
# round-2 GPU call 46 (1 GPU): plain-product knob sweep (in-flight tiles, PDL, tile cuts)
cd "$(dirname "$0")/.."
for env in "X=0" "FPSB_INFLIGHT=3" "FPSB_INFLIGHT=4" "FPSB_NO_PDL=1" "FPSB_NO_CUTS=1" "FPSB_INFLIGHT=1"; do
echo "$env: $(env $env timeout 200 python tools/spmv_bench.py 40 2>&1 | tail -1)"
done
