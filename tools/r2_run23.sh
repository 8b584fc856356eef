#!/usr/bin/env bash
# round-2 GPU call 23 (1 GPU): early row sums (FPSB_LOOP=2) at n = 500 000: which knob / size matters
cd "$(dirname "$0")/.."
run() { env "$@" LMP_REPS=3 FPSB_LOOP=2 timeout 200 python tools/loop_modes_parity.py --size $SZ --tag x 2>&1 | tail -1 | cut -c1-260; }
SZ=500000; run FPSB_LOOP_NSPEC=0
SZ=500000; run FPSB_LOOP_NSPEC=1
SZ=500000; run FPSB_LOOP_NSTAGE=4
SZ=500000; run FPSB_LOOP_CHUNK=100
SZ=500000; run FPSB_LOOP_CHUNK=4
for SZ in 300000 400000 450000 550000 600000 800000; do run FPSB_X=0; done
