#!/usr/bin/env bash
# round-2 GPU call 10 (2 GPUs): world-2 tests of the row-partitioned path + bench.py --gpus 2 with the partitioned record
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2_10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_10_tests.log
tail -3 gpurun_out/r2_10_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_10_bench2.json 2> gpurun_out/r2_10_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_10_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_10_bench2.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1))
for p in d.get("partitioned",[]): print(json.dumps(p))
PY
