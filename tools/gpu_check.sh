#!/bin/bash
# Runs the GPU parity suite in separate processes (a hung kernel then only costs its own group).
# Usage (on the GPU box, from the repo root): bash tools/gpu_check.sh
mkdir -p gpurun_out
run() {  # name, timeout, pytest -k expression
  echo "=== $1" | tee -a gpurun_out/check.log
  timeout -s KILL "$2" python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$3" -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/check.log
  echo "exit=$?" >> gpurun_out/check.log
}
: > gpurun_out/check.log
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv >> gpurun_out/check.log 2>&1
run losses 300 "loss"
run block_fp32 300 "fp32"
run block_bf16_direct 300 "bf16_matches_reference_golden and direct"
run block_bf16_plain_load 300 "bf16_matches_reference_golden and tma_plain_load"
run block_bf16_plain_store 300 "bf16_matches_reference_golden and tma_plain_store"
run block_bf16_tma 300 "bf16_matches_reference_golden and tma and not plain"
run nchw 600 "nchw"
run gemm 400 "gemm or epilogue"
run seeded 600 "seeded"
run full_size 600 "full_size or errors"
echo "=== student" >> gpurun_out/check.log; timeout -s KILL 300 python -m pytest tests/test_student_trainer.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5 >> gpurun_out/check.log
echo "=== smoke" >> gpurun_out/check.log; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -3 >> gpurun_out/check.log
grep -E "^===|passed|failed|exit=|Error|error" gpurun_out/check.log | head -80
