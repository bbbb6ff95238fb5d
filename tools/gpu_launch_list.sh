#!/bin/bash
# the contract's ncu launch list for the default bench workload (durations of every launch of 3 timed steps)
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0"
python bench.py $A > gpurun_out/plain_a.json 2> gpurun_out/plain_a.err &&
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py $A > gpurun_out/ncu_launches.out 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
