#!/usr/bin/env bash
# round-2 GPU call 14: flat one-wide leaf kernels + staged fronts; LDLt tests (incl. BASELINE sizes) and timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ldlt.py tests/test_gpu_baseline_sizes.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2_14_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2_14_tests.log
FPSB_LDLT_MODE=2 timeout 900 python -m pytest tests/test_gpu_ldlt.py -m gpu -x -q 2>&1 | tail -2
for mode in 1 2; do
  echo "== mode $mode"
  FPSB_LDLT_MODE=$mode timeout 600 python tools/ldlt_bench.py 3 2>&1 | tail -3
done
