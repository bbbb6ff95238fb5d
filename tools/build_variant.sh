#!/bin/bash
# A/B build: recompile ONE source with extra -D flags and link it with the other objects of the last build into
# knowledge-distillation-by-replacing-cheap-conv_b200/libkdcc_<tag>.so (select it with KDCC_LIB=...).
#   tools/build_variant.sh w4o3 dw_nhwc3.cu -DKDCC_N3_WARPS=4 -DKDCC_N3_OCC=3
set -e
tag=$1; src=$2; shift 2
P=knowledge-distillation-by-replacing-cheap-conv_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -c $P/csrc/$src -o /tmp/variant_$tag.o
objs=$(ls $P/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $P/libkdcc_$tag.so $objs /tmp/variant_$tag.o -cudart static -Xlinker --no-undefined -ldl -lpthread -lrt
echo built $P/libkdcc_$tag.so
