#!/usr/bin/env bash
# round-2 GPU call 9: full GPU test suite (incl. BASELINE-size parity), full bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_9_tests.log
tail -15 gpurun_out/r2_9_tests.log
timeout 600 python bench.py > gpurun_out/r2_9_bench.json 2> gpurun_out/r2_9_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_9_bench.json")); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3))
print({k:(round(v["us"],1) if "us" in v else round(v.get("ms",0),2)) for k,v in d["extra"].items() if isinstance(v,dict)})
PY
