#!/usr/bin/env bash
# round-2 GPU call 19 (2 GPUs): in-kernel exchange at the full headline size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py $ARGS 2>&1 | grep '^{\|Error\|error' | cut -c1-700 ; }
ARGS="--size 1000000 --tag full_converge"; run FPSB_X=0
ARGS="--size 1000000 --fixed 40 --tag full_fixed40"; run FPSB_X=0
ARGS="--size 1000000 --tag full_converge_launchpath"; run FPSB_DIST_LOOP=0
ARGS="--size 1000000 --tag full_converge_noearly"; run FPSB_LOOP=1
ARGS="--size 500000 --tag half_converge"; run FPSB_X=0
