#!/bin/bash
# Round-2 call 4: v8 attention kernel (four softmax warpgroups): parity + timing + traces with elimination flags
set -u
mkdir -p gpurun_out
echo "=== v8 quick"
timeout 100 python -m pytest tests/test_attention_tc_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2c4_quick.log
if [ "${PIPESTATUS[0]}" = "124" ]; then echo "HANG in the quick test - stopping"; exit 1; fi
echo "=== v8 pytest"
( timeout 400 python -m pytest tests/test_attention_tc_gpu.py tests/test_attention_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -15 ) | tee gpurun_out/r2c4_pytest.log
for lib in "" gpurun_variants/liblongspec_poly0.so gpurun_variants/liblongspec_poly4.so; do
    tag=$([ -z "$lib" ] && echo v8poly2 || basename $lib .so)
    echo "=== $tag"
    ( LONGSPEC_LIB=$lib timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c4_${tag}_verify128k.txt
    ( LONGSPEC_LIB=$lib timeout 60 python tools/prof_attn.py 16384 30 ) 2>&1 | tail -3 | tee gpurun_out/r2c4_${tag}_verify16k.txt
    ( LONGSPEC_LIB=$lib QV=16 timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c4_${tag}_cross.txt
done
for dbg in 0 4; do
    echo "=== trace DBG=$dbg"
    ( DBG=$dbg ROWS=4 timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | tail -14 | tee gpurun_out/r2c4_trace_dbg$dbg.txt
done
