#!/bin/bash
set -u
mkdir -p gpurun_out
for dbg in 0 32 128 160; do
    echo "=== trace DBG=$dbg"
    ( DBG=$dbg ROWS=2 timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | tail -12 | tee gpurun_out/r2c7_trace_dbg$dbg.txt
done
