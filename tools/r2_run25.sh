#!/usr/bin/env bash
# round-2 GPU call 25 (2 GPUs): cheaper ring gate; early row sums on the row-partitioned path; distributed extras at world 2; bench --gpus 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py $ARGS 2>&1 | grep '^{\|Error\|error' | cut -c1-600 ; }
ARGS="--size 1000000 --tag full_converge_early"; run FPSB_DIST_EARLY=1
ARGS="--size 500000 --tag half_converge_early"; run FPSB_DIST_EARLY=1
ARGS="--size 500000 --tag half_converge"; run FPSB_X=0
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_baseline_sizes.py tests/test_gpu_krylov.py -m gpu -x -q 2>&1 | tail -3
for e in 0 1; do
FPSB_DIST_EARLY=$e timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_25_bench2_early$e.json 2> gpurun_out/r2_25_bench2_early$e.err; echo "bench early=$e rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_25_bench2_early$e.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "avg_us", round(d["roofline"]["avg_launch_us"],2))
for p in d.get("partitioned",[]): print(p.get("workload","")[:30], "us/it", round(p.get("us_per_iteration",0),1), "single", round(p.get("single_gpu",{}).get("us_per_iteration",0),1), "speedup", round(p.get("speedup_vs_single_gpu",0),3), "parity", p.get("parity",{}).get("max_rel_err"), p.get("parity",{}).get("ok"), p.get("error"))
PY
done
