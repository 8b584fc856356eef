#!/usr/bin/env bash
# round-2 GPU call 36 (2 GPUs): exchange with the helper CTAs for the boundary rows: tests, parity, bench --gpus 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multiseg.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py $ARGS 2>&1 | grep '^{\|Error\|error' | cut -c1-400 ; }
ARGS="--size 1000000 --tag full_converge"; run FPSB_X=0
ARGS="--size 300000 --fixed 30 --tag fixed"; run FPSB_X=0
FPSB200_LIB=$PWD/variants/libfpsb200_xt.so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/xchg_timers.py --grid 2048 2>&1 | grep '^{' | tee gpurun_out/r2_36_xchg_timers_2gpu.jsonl | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_36_bench2.json 2> gpurun_out/r2_36_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_36_bench2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_36_bench2.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "avg_us", round(d["roofline"]["avg_launch_us"],2), d["clocks"])
for p in d.get("partitioned",[]): print(p.get("workload","")[:30], "us/it", round(p.get("us_per_iteration",0),1), "single", round(p.get("single_gpu",{}).get("us_per_iteration",0),1), "speedup", round(p.get("speedup_vs_single_gpu",0),3), "parity", p.get("parity",{}).get("max_rel_err"), p.get("parity",{}).get("ok"), p.get("error"))
PY
