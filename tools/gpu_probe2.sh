#!/bin/bash
# Runs tools/tc_probe2 over the MMA shapes / TMEM / shuffle questions of the round-2 depthwise redesign (one process per case).
mkdir -p gpurun_out; L=gpurun_out/probe2.log; : > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
r() { timeout -s KILL 60 tools/tc_probe2 "$@" >> $L 2>&1 || echo "FAILED: $*" >> $L; }
for n in 32 48 64 80 96 112 128 160 256; do r ss 128 $n; done
for n in 32 64; do r ssuse 128 $n; done
for n in 32 64 96 128 256; do r ts 128 $n; done
for n in 32 64 128 256; do r ss 64 $n; done
for n in 32 128 256; do r ts 64 $n; done
for m in 32 64 128; do for n in 64 128 256; do r wsss $m $n; r wsts $m $n; done; done
for w in 4 8 16; do r ldtm 0 0 $w; done
for w in 4 8; do r sttm 0 0 $w; done
for w in 4 8 16; do r shfl 0 0 $w; done
cat $L
