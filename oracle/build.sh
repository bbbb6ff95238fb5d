#!/bin/sh
# Builds the CPU oracle (test infrastructure, see kdcc_oracle.c header).
# The reference is pure Python: nothing under /root/reference compiles, so
# there is no oracle/_ref target; reference outputs are frozen as fixtures by make_golden.py.
set -e
cd "$(dirname "$0")"
gcc -O2 -fPIC -fopenmp -fvisibility=hidden -std=c11 -Wall -Wno-comment -Wno-sign-compare \
    -shared -o libkdcc_oracle.so kdcc_oracle.c -lm
