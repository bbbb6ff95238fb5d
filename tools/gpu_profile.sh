#!/bin/bash
# ncu evidence for the default bench workload (one B200).  Each ncu pass only after the same command exited 0 without ncu.
#   1. launch list (duration of every launch) of a short bench run                      -> gpurun_out/launches.csv
#   2. DRAM / L2 / tensor counters of every kdcc kernel of one timed step (few passes)   -> gpurun_out/step_metrics.csv
#   3. --set full + source of the three dominant kernels at the largest site (4096 ch)  -> gpurun_out/prof_*.ncu-rep
mkdir -p gpurun_out
K='dw_tc|pw_gemm|loss|cast_f32|reduce_splits|wgrad2_reduce'
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size'
A="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0"
B="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
python bench.py $A > gpurun_out/plain_a.json 2> gpurun_out/plain_a.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py $A > gpurun_out/ncu_launches.out 2>&1
echo "launch list rc=$?"
python bench.py $B > gpurun_out/plain_b.json 2> gpurun_out/plain_b.err || exit 1
ncu --metrics $M --clock-control none -k regex:"$K" -s 267 -c 89 --csv --log-file gpurun_out/step_metrics.csv python bench.py $B > gpurun_out/ncu_step.out 2>&1
echo "step metrics rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dw_tc_conv2 -s 70 -c 2 -o gpurun_out/prof_conv2 -f python bench.py $B > gpurun_out/ncu_full1.out 2>&1
echo "full conv2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dw_tc_wgrad2_kernel -s 35 -c 1 -o gpurun_out/prof_wgrad2 -f python bench.py $B > gpurun_out/ncu_full2.out 2>&1
echo "full wgrad2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 105 -c 3 -o gpurun_out/prof_gemm -f python bench.py $B > gpurun_out/ncu_full3.out 2>&1
echo "full gemm rc=$?"
ls -la gpurun_out | grep -E "ncu-rep|csv"
