#!/bin/bash
# v7b: prologue overlap + coalesced epilogue: parity, timing, trace; then the whole GPU suite and the bench golden
set -u
mkdir -p gpurun_out
echo "=== quick"
timeout 100 python -m pytest tests/test_attention_tc_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2c9_quick.log
if [ "${PIPESTATUS[0]}" = "124" ]; then echo "HANG in the quick test - stopping"; exit 1; fi
( timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c9_verify128k.txt
( timeout 60 python tools/prof_attn.py 16384 30 ) 2>&1 | tail -3 | tee gpurun_out/r2c9_verify16k.txt
( QV=16 timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c9_cross.txt
for bx in 0 12; do
    ( BX=$bx ROWS=1 timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | grep -E "CTA bx|pair period|pair issued|next P_FULL|S loaded" | tee gpurun_out/r2c9_trace_bx$bx.txt
done
echo "=== full GPU suite"
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/r2c9_pytest_gpu.log
echo "=== bench golden"
( timeout 600 python tools/make_bench_golden.py ) 2>&1 | tail -4 | tee gpurun_out/r2c9_golden.log
