#!/usr/bin/env bash
# round-2 GPU call 13: LDLt with shared-memory staged fronts (FPSB_LDLT_MODE 0 scalar / 1 staged / 2 staged + DMMA)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in 1 2; do
  FPSB_LDLT_MODE=$mode timeout 900 python -m pytest tests/test_gpu_ldlt.py -m gpu -x -q > gpurun_out/r2_13_tests_mode$mode.log 2>&1; echo "mode $mode tests rc=$?"
  tail -3 gpurun_out/r2_13_tests_mode$mode.log
done
for mode in 0 1 2; do
  echo "== mode $mode"
  FPSB_LDLT_MODE=$mode FPSB200_LIB=$PWD/variants/libfpsb200_ldt.so timeout 600 python tools/ldlt_levels.py 2>&1 | tail -5 | cut -c1-1200
done
