#!/usr/bin/env bash
cd "$(dirname "$0")/.."
B="--steps 10 --warmup 3 --no-ldlt --no-cpu-baseline"
run() {
  env "$@" timeout 300 python bench.py $B 2>gpurun_out/r2_8.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print('  value',round(d['value'],1),'avg_us',round(r['avg_launch_us'],2),'frac',round(r['frac'],3), {k:(round(v['us'],1) if 'us' in v else round(v['ms'],2)) for k,v in d['extra'].items() if isinstance(v,dict)})
except Exception as e: print('  failed',e)"
  tail -1 gpurun_out/r2_8.err | cut -c1-300
}
echo "kB=2"; run FPSB_LOOP=2
echo "kB=2 non-early"; run FPSB_LOOP=1
echo "kB=3"; run FPSB200_LIB=$PWD/variants/libfpsb200_kb3.so FPSB_LOOP=2
echo "kB=2 step kernel path"; run FPSB_LOOP=0
FPSB200_LIB=$PWD/variants/libfpsb200_lt.so timeout 300 python tools/loop_timers.py | tail -18
