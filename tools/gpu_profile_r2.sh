#!/bin/bash
# Round-2 ncu evidence for the default bench workload (one B200).  Each ncu pass only after the same command exited 0 without ncu.
#   1. launch list (duration of every launch) of a short bench run                      -> gpurun_out/r2_launches.csv
#   2. DRAM / L2 / tensor counters of every kdcc kernel of one timed step                -> gpurun_out/r2_step_metrics.csv
#   3. --set full + source of the dominant kernels at the largest site (4096 channels)   -> gpurun_out/r2_prof_*.ncu-rep
# bench.py brackets its timed region with cudaProfilerStart/Stop: with --profile-from-start off ncu sees exactly the timed
# step(s), whatever ran before.  Launch order of one step (bench.py --order reference: all forwards, hints, then the backward in
# reverse site order): conv index 0..8 forward (6..8 = the 4096-channel sites), 9..16 dX (9..11 = 4096 channels); weight
# gradient index 0..2 = 4096 channels; GEMM index 6..8 forward of the 4096 -> 256 sites, 9..14 their dX / dW.  FULL=1 runs
# only the three --set full captures.
mkdir -p gpurun_out
K='dw_tc|pw_gemm|loss|cast_f32|reduce_splits|wgrad2_reduce'
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size'
X="--no-cpu-baseline --no-gpu-baseline --no-extras --e2e-steps 0"
A="--steps 3 --warmup 3 $X"
B="--steps 1 --warmup 3 $X"
if [ -z "$FULL" ]; then
python bench.py $A > gpurun_out/r2_plain_a.json 2> gpurun_out/r2_plain_a.err &&
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py $A > gpurun_out/r2_ncu_launches.out 2>&1
echo "launch list rc=$?"
python bench.py $B > gpurun_out/r2_plain_b.json 2> gpurun_out/r2_plain_b.err || exit 1
timeout -s KILL 600 ncu --metrics $M --clock-control none --profile-from-start off -k regex:"$K" --csv --log-file gpurun_out/r2_step_metrics.csv python bench.py $B > gpurun_out/r2_ncu_step.out 2>&1
echo "step metrics rc=$?"
else python bench.py $B > gpurun_out/r2_plain_b.json 2> gpurun_out/r2_plain_b.err || exit 1; fi
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dw_tc_conv2 -s 8 -c 2 -o gpurun_out/r2_prof_conv2 -f python bench.py $B > gpurun_out/r2_ncu_full1.out 2>&1
echo "full conv2 rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dw_tc_wgrad3_kernel -s 0 -c 1 -o gpurun_out/r2_prof_wgrad3 -f python bench.py $B > gpurun_out/r2_ncu_full2.out 2>&1
echo "full wgrad3 rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pw_gemm -s 8 -c 3 -o gpurun_out/r2_prof_gemm -f python bench.py $B > gpurun_out/r2_ncu_full3.out 2>&1
echo "full gemm rc=$?"
[ -z "$FULL" ] && python tools/step_metrics_summary.py gpurun_out/r2_step_metrics.csv gpurun_out/r2_step_metrics.txt gpurun_out/r2_traffic.json | tail -20
for f in conv2 wgrad3 gemm; do python tools/ncu_summary.py gpurun_out/r2_prof_$f.ncu-rep > gpurun_out/r2_ncu_$f.txt 2>&1; done
ls -la gpurun_out | grep -E "r2_.*(ncu-rep|csv|txt|json)"
