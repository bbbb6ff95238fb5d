#!/bin/bash
# Bench + ncu evidence on one B200 (run from the repo root on the GPU box).
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err
A="--steps 2 --warmup 3 --batch 2 --no-cpu-baseline --e2e-steps 0"
python bench.py $A > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py $A > gpurun_out/ncu_launches.out 2>&1
echo "launch list rc=$?"
B="--steps 1 --warmup 3 --batch 1 --no-cpu-baseline --e2e-steps 0"
python bench.py $B > gpurun_out/plain_b1.json 2> gpurun_out/plain_b1.err &&
ncu --set full --clock-control none --import-source on -k regex:'dw_tma|pw_gemm_sm100|hint_loss_kernel|kd_loss_kernel' -s 330 -c 40 -o gpurun_out/prof_r01 -f python bench.py $B > gpurun_out/ncu_full.out 2>&1
echo "full rc=$?"
ls -la gpurun_out
