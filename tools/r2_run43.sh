#!/usr/bin/env bash
# round-2 GPU call 43 (1 GPU): the driver's own sequence on the final tree: GPU tests, smoke, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_43_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_43_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_43_bench.json 2> gpurun_out/r2_43_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_43_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3), d["clocks"], d["gpu_launches"])
print({k:(round(v["us"],1) if "us" in v else round(v.get("ms",0),2)) for k,v in d["extra"].items() if isinstance(v,dict)})
PY
