K='dw_tc|pw_gemm|loss|cast_f32|reduce_splits|wgrad2_reduce'
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size'
B="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
python bench.py $B > gpurun_out/plain_b.json 2> gpurun_out/plain_b.err || exit 1
ncu --metrics $M --clock-control none -k regex:"$K" -s 267 -c 89 --csv --log-file gpurun_out/step_metrics.csv python bench.py $B > gpurun_out/ncu_step.out 2>&1
echo "step metrics rc=$?"
