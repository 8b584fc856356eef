#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_3_tests.log
tail -4 gpurun_out/r2_3_tests.log
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
for v in 0 2; do
  FPSB_LOOP=$v timeout 300 python bench.py $B > gpurun_out/r2_3_bench_loop$v.json 2> gpurun_out/r2_3_bench_loop$v.err
  echo "loop=$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_3_bench_loop$v.json"))
    r=d["roofline"]
    print("  value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3),"iters",r["iters"],"extra",{k:(round(v["us"],1) if "us" in v else round(v["ms"],2)) for k,v in d["extra"].items() if isinstance(v,dict)})
except Exception as e:
    print("  parse failed",e)
PY
done
FPSB200_LIB=$PWD/variants/libfpsb200_lt.so timeout 300 python tools/loop_timers.py
