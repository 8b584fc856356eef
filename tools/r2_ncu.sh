#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# one full-set capture of the first persistent-loop launch of the third solve
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gk_loop_kernel -s 6 -c 1 -o gpurun_out/r2_loop_full -f python tools/solve_bench.py 4 > gpurun_out/r2_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu.log
ls -la gpurun_out/r2_loop_full.ncu-rep
