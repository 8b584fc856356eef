#!/usr/bin/env bash
# round-2 GPU call 48 (1 GPU): bench with per-call LDLt extras
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_48_bench.json 2> gpurun_out/r2_48_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_48_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2))
print(d["extra"]["ldlt_solve_two_mixed"], d["extra"]["ldlt_solve_two_least_squares"])
PY
