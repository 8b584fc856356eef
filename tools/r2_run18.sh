#!/usr/bin/env bash
# round-2 GPU call 18 (2 GPUs): what breaks the in-kernel exchange on the headline operator
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py --size 200000 $ARGS 2>&1 | grep '^{' ; }
ARGS="--fixed 30 --tag fixed30"; run FPSB_X=0
ARGS="--fixed 30 --tag fixed30_launchpath"; run FPSB_DIST_LOOP=0
ARGS="--fixed 30 --tag fixed30_noearly"; run FPSB_LOOP=1
ARGS="--fixed 30 --tag fixed30_nspec0"; run FPSB_LOOP_NSPEC=0
ARGS="--fixed 30 --tag fixed30_chunk100"; run FPSB_LOOP_CHUNK=100
ARGS="--fixed 30 --tag fixed30_chunk1"; run FPSB_LOOP_CHUNK=1
ARGS="--tag converge"; run FPSB_X=0
ARGS="--fixed 30 --delta 0.01 --tag fixed30_delta"; run FPSB_X=0
