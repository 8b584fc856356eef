#!/usr/bin/env bash
# round-2 GPU call 17: ncu evidence. (a) LDLt refactor + solve per kernel, FP64 pipe vs DMMA pipe (FPSB_LDLT_MODE 1 / 2);
# (b) launch list of the default bench command; (c) one full-set capture of the persistent loop kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
for mode in 1 2; do
  FPSB_LDLT_MODE=$mode timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_17_ldlt_mode$mode.csv python tools/ldlt_bench.py 1 > gpurun_out/r2_17_ldlt_mode$mode.log 2>&1
  echo "ncu ldlt mode $mode rc=$?"; tail -1 gpurun_out/r2_17_ldlt_mode$mode.log
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_17_bench_plain.json 2> gpurun_out/r2_17_bench_plain.err; echo "plain bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_17_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_17_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gk_loop_kernel -s 6 -c 1 -o gpurun_out/r2_17_loop_full -f python tools/solve_bench.py 4 > gpurun_out/r2_17_ncu_loop.log 2>&1
echo "ncu loop rc=$?"
