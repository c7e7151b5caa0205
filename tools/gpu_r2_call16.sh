#!/bin/bash
# own prefill (tcgen05 causal blocks), seams, unmodified driver: full GPU suite + smoke + bench
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/r2c16_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) | tee gpurun_out/r2c16_smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; echo "bench exit $?" >> gpurun_out/r2c16_bench.err )
tail -3 gpurun_out/r2c16_bench.err | cut -c1-300; cut -c1-900 gpurun_out/r2c16_bench.json
