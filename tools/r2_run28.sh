#!/usr/bin/env bash
# round-2 GPU call 28 (1 GPU): multi-segment gather windows: tests, C3 single-GPU timing with / without, headline sanity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiseg.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/poisson_time.py 2048 2>&1 | tail -2
FPSB_NO_SEGS=1 timeout 300 python tools/poisson_time.py 2048 2>&1 | tail -2
timeout 300 python tools/poisson_time.py 1024 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_28_bench.json 2> gpurun_out/r2_28_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_28_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "avg_us", round(d["roofline"]["avg_launch_us"],2))
print({k:(round(v.get("us",0),1) if isinstance(v,dict) else v) for k,v in d["extra"].items() if k.startswith("spmv")})
PY
