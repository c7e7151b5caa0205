#!/bin/bash
# Short GPU session: attention parity tests + per-launch timings of the verify / draft-cross shapes.
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -15 gpurun_out/pytest_gpu.log
( timeout 200 python tools/prof_attn.py 131072 8 ) 2>&1 | tee gpurun_out/prof_attn_128k.txt
( QV=16 timeout 200 python tools/prof_attn.py 131072 8 ) 2>&1 | tee gpurun_out/prof_attn_128k_qv16.txt
( timeout 200 python tools/prof_attn.py 16384 8 ) 2>&1 | tee gpurun_out/prof_attn_16k.txt
