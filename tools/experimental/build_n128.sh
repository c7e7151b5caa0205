#!/bin/bash
# Builds tools/experimental/liblongspec_n128.so: the product sources with attn_tc.cu replaced by the experimental
# N = 128 shared-accumulator variant.  Try it with:  LONGSPEC_LIB=tools/experimental/liblongspec_n128.so pytest tests -m gpu
set -e
cd "$(dirname "$0")/../.."
SRC=$(ls longspec_b200/csrc/*.cu | grep -v attn_tc.cu)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC ${EXTRA:-} \
    -o tools/experimental/liblongspec_n128.so $SRC tools/experimental/attn_tc_n128.cu -lcuda
echo built tools/experimental/liblongspec_n128.so
