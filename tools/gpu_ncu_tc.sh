#!/bin/bash
mkdir -p gpurun_out
B="--steps 1 --warmup 3 --batch 1 --no-cpu-baseline --e2e-steps 0"
python bench.py $B > gpurun_out/plain_b1.json 2> gpurun_out/plain_b1.err &&
ncu --set full --clock-control none --import-source on -k regex:dw_tc -s 81 -c 27 -o gpurun_out/prof_dwtc_r01 -f python bench.py $B > gpurun_out/ncu_full.out 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.out | cut -c1-300
