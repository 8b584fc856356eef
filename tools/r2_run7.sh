#!/usr/bin/env bash
cd "$(dirname "$0")/.."
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
run() {
  env "$@" timeout 300 python bench.py $B 2>gpurun_out/r2_7.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print('  value',round(d['value'],1),'avg_us',round(r['avg_launch_us'],2),'frac',round(r['frac'],3))
except Exception as e: print('  failed',e)"
  tail -1 gpurun_out/r2_7.err | cut -c1-300
}
for cfg in "FPSB_LOOP=2 FPSB_LOOP_ODD=1" "FPSB_LOOP=2 FPSB_LOOP_ODD=1 FPSB_NO_CUTS=1" "FPSB_LOOP=2 FPSB_LOOP_CHUNK=48" "FPSB_LOOP=0"; do
  echo "$cfg"; run $cfg
done
