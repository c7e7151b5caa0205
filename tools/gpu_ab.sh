#!/bin/bash
# A/B of the production library against an alternative build of the same ABI (default: the experimental N=128 variant):
# attention parity tests + per-launch timing of the verify / draft-cross shapes with each.
#   gpurun --timeout 600 -- 'bash tools/gpu_ab.sh [path/to/alt.so]'
set -u
mkdir -p gpurun_out
ALT=${1:-tools/experimental/liblongspec_n128.so}
for lib in "" "$ALT"; do
    tag=$([ -z "$lib" ] && echo prod || echo alt)
    echo "=== $tag ${lib}"
    ( LONGSPEC_LIB=$lib timeout 300 python -m pytest tests/test_attention_tc_gpu.py tests/test_attention_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/ab_${tag}_pytest.log
    ( LONGSPEC_LIB=$lib timeout 120 python tools/prof_attn.py 131072 12 ) 2>&1 | tee gpurun_out/ab_${tag}_verify.txt
    ( LONGSPEC_LIB=$lib QV=16 timeout 120 python tools/prof_attn.py 131072 12 ) 2>&1 | tee gpurun_out/ab_${tag}_cross.txt
done
