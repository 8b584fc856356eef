#!/usr/bin/env bash
# round-2 GPU call 44 (1 GPU): pipelined e2e with / without the compute gate (same box, alternating, two runs each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
for G in gate nogate; do
if [ $G = nogate ]; then export FPSB_BENCH_NO_GATE=1; else unset FPSB_BENCH_NO_GATE; fi
timeout 600 python bench.py --no-cpu-baseline --no-ldlt > gpurun_out/r2_44_bench_${G}_$i.json 2> gpurun_out/r2_44_bench_${G}_$i.err; echo "bench $G $i rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_44_bench_${G}_$i.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("$G $i value",round(d["value"],1),"e2e",round(e["value"],1),round(e["ms_per_step"],2),"serial",round(e["one_at_a_time"]["value"],1))
PY
done; done
