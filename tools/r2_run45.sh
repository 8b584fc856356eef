#!/usr/bin/env bash
# round-2 GPU call 45 (2 GPUs): bench.py --gpus 2 of the final tree (compute gate under torchrun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_45_bench2.json 2> gpurun_out/r2_45_bench2.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_45_bench2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_45_bench2.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "avg_us", round(d["roofline"]["avg_launch_us"],2), d["clocks"])
for p in d.get("partitioned",[]): print(p.get("workload","")[:30], "us/it", round(p.get("us_per_iteration",0),1), "single", round(p.get("single_gpu",{}).get("us_per_iteration",0),1), "speedup", round(p.get("speedup_vs_single_gpu",0),3), "parity", p.get("parity",{}).get("max_rel_err"), p.get("parity",{}).get("ok"), p.get("error"))
PY
