#!/bin/bash
# hardware probes that bound the streaming NHWC 3x3 kernel (DESIGN.md 4.1 c'): fp32 pipe rates and HBM bandwidth by run length
for t in fma_probe dram_chunk_probe; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/$t tools/$t.cu && timeout -s KILL 120 tools/$t
done
