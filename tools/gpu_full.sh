#!/bin/bash
# Full verification as the driver would run it: pytest -m gpu (whole suite, one process), smoke, default bench, reference arm.
mkdir -p gpurun_out
bash tools/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed|smoke" gpurun_out/check.log
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"
