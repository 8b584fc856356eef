#!/usr/bin/env bash
# round-2 GPU call 30 (1 GPU): device-resident Val(1) hprod, fused feasibility step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_nlp.py tests/test_gpu_feasibility.py -m gpu -q 2>&1 | tail -40
