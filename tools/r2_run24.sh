#!/usr/bin/env bash
# round-2 GPU call 24 (1 GPU): ring release counters (stage_turn_wait): early row sums at the sizes that failed, full GPU test suite, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { env "$@" LMP_REPS=4 timeout 200 python tools/loop_modes_parity.py --size $SZ --tag $TAG 2>&1 | tail -1 | cut -c1-260; }
for SZ in 450000 500000 550000; do TAG=e$SZ; run FPSB_LOOP=2; done
SZ=500000; TAG=m0; run FPSB_LOOP=0
python tools/loop_modes_parity.py --compare e500000 m0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_24_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_24_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_24_bench.json 2> gpurun_out/r2_24_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_24_bench.json")); r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",round(d["e2e"]["one_at_a_time"]["value"],1),"avg_us",round(r["avg_launch_us"],2),"frac",round(r["frac"],3))
print({k:(round(v["us"],1) if "us" in v else round(v.get("ms",0),2)) for k,v in d["extra"].items() if isinstance(v,dict)})
PY
