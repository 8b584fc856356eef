#!/usr/bin/env bash
cd "$(dirname "$0")/.."
FPSB200_LIB=$PWD/variants/libfpsb200_ldt.so timeout 600 python tools/ldlt_levels.py 2>&1 | tail -8
timeout 300 python bench.py --no-ldlt --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v['us'],1) if 'us' in v else round(v.get('ms',0),2)) for k,v in d['extra'].items() if isinstance(v,dict)})"
