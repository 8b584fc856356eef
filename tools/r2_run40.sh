#!/usr/bin/env bash
# round-2 GPU call 40 (1 GPU): explicit linear constraints on the GPU solvers, bench sanity (gc / median changes), ncu --set full of the
# persistent loop kernel on the C3 operator with multi-segment windows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fps_solve.py tests/test_gpu_trcg.py tests/test_gpu_feasibility.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline --no-ldlt > gpurun_out/r2_40_bench.json 2> gpurun_out/r2_40_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_40_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3), d["clocks"])
print(d["extra"].get("iter_solve_two_least_squares"), {k:round(v["us"],1) for k,v in d["extra"].items() if k.startswith("spmv")})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gk_loop_kernel -s 4 -c 1 -o gpurun_out/r2_40_loop_c3_full -f python tools/poisson_time.py 2048 > gpurun_out/r2_40_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_40_ncu.log; ls -la gpurun_out/r2_40_loop_c3_full.ncu-rep
