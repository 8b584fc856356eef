#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
for v in "2 0" "1 0" "2 1" "1 1"; do
  set -- $v
  FPSB_LOOP=$1 FPSB_LOOP_RPLACE=$2 timeout 300 python bench.py $B > gpurun_out/r2_5_bench_$1_$2.json 2> gpurun_out/r2_5_bench_$1_$2.err
  echo "loop=$1 rplace=$2 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_5_bench_$1_$2.json"))
    r=d["roofline"]
    print("  value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3),"iters",r["iters"])
except Exception as e:
    print("  parse failed",e)
PY
done
