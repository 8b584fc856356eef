#!/usr/bin/env bash
# round-2 GPU call 21 (1 GPU): do the three single-GPU loop modes agree at n = 500 000 / 1 000 000?  bench extras anomaly
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sz in 500000 1000000; do
  for md in 0 1 2; do FPSB_LOOP=$md timeout 300 python tools/loop_modes_parity.py --size $sz --tag s${sz}_m$md 2>&1 | tail -1; done
  python tools/loop_modes_parity.py --compare s${sz}_m1 s${sz}_m0
  python tools/loop_modes_parity.py --compare s${sz}_m2 s${sz}_m0
done
timeout 600 python bench.py --no-ldlt --no-cpu-baseline --e2e-serial --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('serial-e2e bench extras', {k:(round(v['us'],1) if 'us' in v else round(v.get('ms',0),2)) for k,v in d['extra'].items() if isinstance(v,dict)})"
timeout 600 python bench.py --no-ldlt --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pipelined-e2e bench extras', {k:(round(v['us'],1) if 'us' in v else round(v.get('ms',0),2)) for k,v in d['extra'].items() if isinstance(v,dict)}, 'value', d['value'], 'e2e', d['e2e']['value'])"
