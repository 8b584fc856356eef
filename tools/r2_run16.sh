#!/usr/bin/env bash
# round-2 GPU call 16 (2 GPUs): bench.py --gpus 2 with the in-kernel exchange (partitioned record)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-ldlt --no-cpu-baseline > gpurun_out/r2_16_bench2.json 2> gpurun_out/r2_16_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_16_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_16_bench2.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1))
for p in d.get("partitioned",[]): print(json.dumps({k:v for k,v in p.items() if k not in ("kernel","path")}))
PY
