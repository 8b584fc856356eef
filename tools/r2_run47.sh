#!/usr/bin/env bash
# round-2 GPU call 47 (1 GPU): default bench of the final tree (product extras over 100 launches), Krylov GPU tests after the in-flight clamp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_multiseg.py tests/test_gpu_baseline_sizes.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_47_bench.json 2> gpurun_out/r2_47_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_47_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3), d["clocks"], d["gpu_launches"])
print({k:(round(v["us"],1), round(v["frac_of_measured_peak"],3)) for k,v in d["extra"].items() if k.startswith("spmv")}, d["extra"]["iter_solve_two_least_squares"]["ms"], d["extra"]["ldlt_solve_two_mixed"]["ms"], d["cpu_baseline"]["value"])
PY
