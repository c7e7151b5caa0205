#!/bin/bash
# ncu --set full of the v6 verify-attention kernel (source-level stall samples) + finer issuer trace
set -u
mkdir -p gpurun_out
( timeout 60 python tools/trace_attn.py 131072 ) 2>&1 | tail -16 | tee gpurun_out/r2c3_trace.txt
( ROWS=2 BX=1 timeout 60 python tools/trace_attn.py 16384 ) 2>&1 | tail -14 | tee gpurun_out/r2c3_trace16k.txt
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 1 -o gpurun_out/r2c3_attn_tc_full -f \
    python tools/prof_attn.py 131072 3 > gpurun_out/r2c3_ncu_full.log 2>&1 )
tail -5 gpurun_out/r2c3_ncu_full.log
ls -la gpurun_out/*.ncu-rep
