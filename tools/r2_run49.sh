#!/usr/bin/env bash
# round-2 GPU call 49 (1 GPU): canonical default bench of the final tree (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_49_bench.json 2> gpurun_out/r2_49_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_49_bench_reference.json 2> gpurun_out/r2_49_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_49_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3), d["clocks"], d["gpu_launches"])
x=d["extra"]; print({k:(round(v["us"],1), round(v["frac_of_measured_peak"],3)) for k,v in x.items() if k.startswith("spmv")}, x["iter_solve_two_least_squares"]["ms"], x["ldlt_solve_two_mixed"]["ms"], x["ldlt_solve_two_mixed"]["ms_all"], x["ldlt_solve_two_least_squares"]["ms"], d["cpu_baseline"]["value"])
r=json.loads(open("gpurun_out/r2_49_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
