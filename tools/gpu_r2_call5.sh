#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 1 -o gpurun_out/r2c5_attn_tc_v7 -f \
    python tools/prof_attn.py 131072 3 > gpurun_out/r2c5_ncu.log 2>&1 )
tail -3 gpurun_out/r2c5_ncu.log
