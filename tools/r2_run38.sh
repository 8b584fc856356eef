#!/usr/bin/env bash
# round-2 GPU call 38 (1 GPU): final validation: full GPU test suite, smoke, bench (both arms), ncu launch list, BASELINE configs C2 / C4 / C5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_38_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_38_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_38_bench.json 2> gpurun_out/r2_38_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_38_bench_reference.json 2> gpurun_out/r2_38_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_38_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",round(d["e2e"]["one_at_a_time"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3), d["clocks"], "cpu", d["cpu_baseline"])
print({k:(round(v["us"],1) if "us" in v else round(v.get("ms",0),2)) for k,v in d["extra"].items() if isinstance(v,dict)})
print(open("gpurun_out/r2_38_bench_reference.json").read()[:600])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_38_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ldlt > gpurun_out/r2_38_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 python tools/configs_bench.py --skip-cpu > gpurun_out/r2_38_configs.jsonl 2> gpurun_out/r2_38_configs.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/r2_38_configs.jsonl
timeout 600 python tools/c4_rankdef.py > gpurun_out/r2_38_c4.jsonl 2> gpurun_out/r2_38_c4.err; echo "c4 rc=$?"; cut -c1-500 gpurun_out/r2_38_c4.jsonl
