#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nchw" -p no:cacheprovider -x 2>&1 | tail -60 > gpurun_out/nchw.log
grep -E "passed|failed|Error|error|assert" gpurun_out/nchw.log | head -30
