#!/usr/bin/env bash
# round-2 GPU call 34 (8 GPUs): segment timers of the in-kernel exchange on the C3 operator
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FPSB200_LIB=$PWD/variants/libfpsb200_xt.so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/xchg_timers.py --grid 2048 2>&1 | grep '^{' | tee gpurun_out/r2_34_xchg_timers_8gpu.jsonl | cut -c1-900
