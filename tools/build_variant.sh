#!/bin/bash
# Alternative builds of the SAME ABI for A/B runs on the GPU box (selected with LONGSPEC_LIB=...):
#   tools/build_variant.sh <name> [extra nvcc flags, e.g. -DLS_TC_POLY=4]   ->  gpurun_variants/liblongspec_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC "$@" \
    -o gpurun_variants/liblongspec_${name}.so longspec_b200/csrc/*.cu -lcuda
echo built gpurun_variants/liblongspec_${name}.so
