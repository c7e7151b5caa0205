#!/bin/bash
# Round-2 call 1: softmax-math micro-benchmark, FA2 (flash-attn wheel) vs ours at 128k / 16k, A/B of the N=128 variant.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
( timeout 120 tools/ubench/softmax_math ) 2>&1 | tee gpurun_out/r2c1_softmax_math.txt
( timeout 300 python tools/ref_gpu_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c1_ref_128k.txt
( timeout 300 python tools/ref_gpu_attn.py 16384 30 ) 2>&1 | tail -3 | tee gpurun_out/r2c1_ref_16k.txt
( QV=16 timeout 300 python tools/ref_gpu_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c1_ref_cross.txt
ALT=tools/experimental/liblongspec_n128.so
echo "=== alt pytest"
( LONGSPEC_LIB=$ALT timeout 400 python -m pytest tests/test_attention_tc_gpu.py tests/test_attention_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -15 ) | tee gpurun_out/r2c1_alt_pytest.log
for lib in "" "$ALT"; do
    tag=$([ -z "$lib" ] && echo prod || echo alt)
    ( LONGSPEC_LIB=$lib timeout 120 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -4 | tee gpurun_out/r2c1_${tag}_verify128k.txt
    ( LONGSPEC_LIB=$lib timeout 120 python tools/prof_attn.py 16384 30 ) 2>&1 | tail -4 | tee gpurun_out/r2c1_${tag}_verify16k.txt
    ( LONGSPEC_LIB=$lib QV=16 timeout 120 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -4 | tee gpurun_out/r2c1_${tag}_cross.txt
done
