#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --layout nchw > gpurun_out/bench_nchw.json 2> gpurun_out/bench_nchw.err
echo "nchw rc=$?"; tail -c 400 gpurun_out/bench_nchw.err
python bench.py --steps 10 --warmup 3 --layout nhwc --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
echo "nhwc rc=$?"
KDCC_DW_TC_SINGLE=0 python bench.py --steps 10 --warmup 3 --layout nchw --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_nchw_pertap.json 2>&1
bash tools/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed" gpurun_out/check.log
