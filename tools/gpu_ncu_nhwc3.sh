#!/bin/bash
mkdir -p gpurun_out
B="--dw k3d1p1 --layout nhwc --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
python bench.py $B > gpurun_out/plain_n3.json 2> gpurun_out/plain_n3.err &&
ncu --set full --clock-control none --import-source on -k regex:dw_nhwc3 -s 105 -c 3 -o gpurun_out/prof_nhwc3 -f python bench.py $B > gpurun_out/ncu_n3.out 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_nhwc3.ncu-rep
