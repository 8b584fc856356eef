#!/usr/bin/env bash
cd "$(dirname "$0")/.."
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
for v in 2 3 4 5; do
  FPSB_INFLIGHT=$v timeout 300 python bench.py $B > gpurun_out/r2_2_inflight$v.json 2> gpurun_out/r2_2_inflight$v.err
  echo "inflight=$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_2_inflight$v.json"))
    r=d["roofline"]
    print("  value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3),"iters",r["iters"])
except Exception as e:
    print("  parse failed",e)
PY
done
FPSB_INFLIGHT=4 FPSB200_LIB=$PWD/variants/libfpsb200_lt.so timeout 300 python tools/loop_timers.py
