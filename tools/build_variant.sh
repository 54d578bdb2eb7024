#!/bin/bash
# Developer build of libb200render with another CTA size for the persistent ray-tracing kernels:
#   tools/build_variant.sh 64   ->  renderer_b200/libb200render_b64.so   (select with B200R_LIB=...)
set -e
B=$1; cd "$(dirname "$0")/.."
O=renderer_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
     --expt-relaxed-constexpr -DB200R_RT_BLOCK=$B -c renderer_b200/csrc/cuda/rt_kernels.cu -o $O/rt_kernels_b$B.o
OBJS=$(ls $O/*.o | grep -v "rt_kernels" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o renderer_b200/libb200render_b$B.so $OBJS $O/rt_kernels_b$B.o -lpthread
echo built renderer_b200/libb200render_b$B.so
