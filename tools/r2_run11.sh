#!/usr/bin/env bash
# round-2 GPU call 11: even gather windows (plain SpMV), pipelined e2e, Krylov tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2_11_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_11_tests.log
tail -3 gpurun_out/r2_11_tests.log
timeout 600 python bench.py --no-ldlt --no-cpu-baseline > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_11_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_11_bench.json")); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",round(d["e2e"]["one_at_a_time"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3))
print({k:(round(v["us"],1) if "us" in v else round(v.get("ms",0),2)) for k,v in d["extra"].items() if isinstance(v,dict)})
PY
