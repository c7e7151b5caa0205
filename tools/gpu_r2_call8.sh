#!/bin/bash
set -u
mkdir -p gpurun_out
for bx in 0 12 17; do
    echo "=== trace BX=$bx 128k"
    ( BX=$bx ROWS=1 timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | tail -13 | tee gpurun_out/r2c8_trace_bx$bx.txt
done
for bx in 0 12; do
    echo "=== trace BX=$bx 16k"
    ( BX=$bx ROWS=1 timeout 60 python tools/trace_attn.py 16384 ) 2>&1 | grep -E "CTA bx|alone|pair period|pair issued|next P_FULL" | tee gpurun_out/r2c8_trace16k_bx$bx.txt
done
