#!/usr/bin/env bash
# round-2 GPU call 20 (2 GPUs): in-kernel exchange without early row sums: parity at full size, world-2 tests, bench --gpus 2; lsq anomaly check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py $ARGS 2>&1 | grep '^{\|Error\|error' | cut -c1-700 ; }
ARGS="--size 1000000 --tag full_converge"; run FPSB_X=0
ARGS="--size 500000 --tag half_converge"; run FPSB_X=0
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_20_bench2.json 2> gpurun_out/r2_20_bench2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_20_bench2.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1))
for p in d.get("partitioned",[]): print(json.dumps({k:v for k,v in p.items() if k not in ("kernel","path","launch_per_half_iteration_path","per_launch_us_max_over_ranks")}))
PY
timeout 300 python tools/lsq_time.py 2>&1 | tail -24
