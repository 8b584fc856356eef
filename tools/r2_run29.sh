#!/usr/bin/env bash
# round-2 GPU call 29 (1 GPU): multi-segment windows with the 1 664-entry cap, device-resident Val(1) hprod
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiseg.py tests/test_gpu_device_nlp.py -m gpu -q 2>&1 | tail -25
timeout 300 python tools/poisson_time.py 2048 2>&1 | tail -2
timeout 300 python tools/poisson_time.py 1024 2>&1 | tail -2
