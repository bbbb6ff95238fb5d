#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_student_trainer.py -q -m gpu -p no:cacheprovider 2>&1 | grep -E "assert|Error|passed|failed" | head -10
