#!/bin/bash
# Round-2 call 2: v6 attention kernel (N=128 pairs, setmaxnreg, polynomial exp2 offload): parity + timing + trace.
set -u
mkdir -p gpurun_out
echo "=== v6 quick"
timeout 100 python -m pytest tests/test_attention_tc_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2c2_quick.log
if [ "${PIPESTATUS[0]}" = "124" ]; then echo "HANG in the quick test - stopping"; exit 1; fi
echo "=== v6 pytest"
( timeout 300 python -m pytest tests/test_attention_tc_gpu.py tests/test_attention_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -15 ) | tee gpurun_out/r2c2_pytest.log
for lib in "" gpurun_variants/liblongspec_poly0.so gpurun_variants/liblongspec_poly4.so gpurun_variants/liblongspec_r1prod.so; do
    tag=$([ -z "$lib" ] && echo v6poly2 || basename $lib .so)
    echo "=== $tag"
    ( LONGSPEC_LIB=$lib timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c2_${tag}_verify128k.txt
    ( LONGSPEC_LIB=$lib timeout 60 python tools/prof_attn.py 16384 30 ) 2>&1 | tail -3 | tee gpurun_out/r2c2_${tag}_verify16k.txt
    ( LONGSPEC_LIB=$lib QV=16 timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c2_${tag}_cross.txt
done
echo "=== trace"
( timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | tail -40 | tee gpurun_out/r2c2_trace.txt
