#!/bin/bash
# v9: two MMA issuers
set -u
mkdir -p gpurun_out
echo "=== quick"
timeout 100 python -m pytest tests/test_attention_tc_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2c10_quick.log
if [ "${PIPESTATUS[0]}" = "124" ]; then echo "HANG in the quick test - stopping"; exit 1; fi
( timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c10_verify128k.txt
( timeout 60 python tools/prof_attn.py 16384 30 ) 2>&1 | tail -3 | tee gpurun_out/r2c10_verify16k.txt
( QV=16 timeout 60 python tools/prof_attn.py 131072 12 ) 2>&1 | tail -3 | tee gpurun_out/r2c10_cross.txt
for bx in 0 12; do
    ( BX=$bx ROWS=1 timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | grep -E "CTA bx|pair period|pair issued|next P_FULL|S loaded|PV issued|K tile" | tee gpurun_out/r2c10_trace_bx$bx.txt
done
( timeout 300 python -m pytest tests/test_attention_tc_gpu.py tests/test_attention_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -5 ) | tee gpurun_out/r2c10_pytest.log
