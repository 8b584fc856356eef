#!/usr/bin/env bash
# round-2 GPU call 22 (1 GPU): early row sums (FPSB_LOOP=2): which solve goes wrong, and why (initcheck / racecheck on a small operator)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do LMP_REPS=5 FPSB_LOOP=2 timeout 300 python tools/loop_modes_parity.py --size 500000 --tag e$i 2>&1 | tail -2 | cut -c1-400; done
LMP_REPS=5 FPSB_LOOP=1 timeout 300 python tools/loop_modes_parity.py --size 500000 --tag ne 2>&1 | tail -2 | cut -c1-400
LMP_REPS=5 FPSB_LOOP=2 timeout 300 python tools/loop_modes_parity.py --size 700000 --tag e700 2>&1 | tail -2 | cut -c1-400
LMP_REPS=1 FPSB_LOOP=2 timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/loop_modes_parity.py --size 60000 --tag ic > gpurun_out/r2_22_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -c "Uninitialized" gpurun_out/r2_22_initcheck.log; grep -A12 "Uninitialized" gpurun_out/r2_22_initcheck.log | head -60 | cut -c1-220; tail -3 gpurun_out/r2_22_initcheck.log | cut -c1-300
LMP_REPS=1 FPSB_LOOP=2 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/loop_modes_parity.py --size 60000 --tag rc > gpurun_out/r2_22_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/r2_22_racecheck.log; grep -B2 -A14 "hazard" gpurun_out/r2_22_racecheck.log | head -80 | cut -c1-220; tail -3 gpurun_out/r2_22_racecheck.log | cut -c1-300
