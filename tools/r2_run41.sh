#!/usr/bin/env bash
# round-2 GPU call 41 (2 GPUs): world-2 tests of the final build incl. the long-boundary (helper CTA) regression test
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -6
