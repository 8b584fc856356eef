#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_6_tests.log
tail -3 gpurun_out/r2_6_tests.log
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
for v in "2 3" "1 3" "2 2"; do
  set -- $v
  FPSB_LOOP=$1 FPSB_LOOP_NSPEC=$2 timeout 300 python bench.py $B > gpurun_out/r2_6_bench_$1_$2.json 2> gpurun_out/r2_6_bench_$1_$2.err
  echo "loop=$1 nspec=$2 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_6_bench_$1_$2.json"))
    r=d["roofline"]
    print("  value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3),"iters",r["iters"])
except Exception as e:
    print("  parse failed",e)
PY
done
timeout 200 python tools/kinds_bench.py
FPSB200_LIB=$PWD/variants/libfpsb200_lt.so timeout 300 python tools/loop_timers.py
