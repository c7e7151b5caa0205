#!/bin/bash
# bench at N=1 (ours + reference arm) and the other BASELINE configs
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c11_gpu.txt 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "bench exit $?" >> gpurun_out/r2c11_bench.err )
tail -3 gpurun_out/r2c11_bench.err; cat gpurun_out/r2c11_bench.json | cut -c1-1500
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2c11_bench_ref.json 2> gpurun_out/r2c11_bench_ref.err; echo "ref exit $?" >> gpurun_out/r2c11_bench_ref.err )
tail -3 gpurun_out/r2c11_bench_ref.err; cat gpurun_out/r2c11_bench_ref.json | cut -c1-800
# config 2: 16k prompt
( LONGSPEC_BENCH_PROMPT=16384 timeout 300 python tools/make_bench_golden.py ) 2>&1 | tail -2 | cut -c1-400
( LONGSPEC_BENCH_PROMPT=16384 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c11_bench_16k.json 2> gpurun_out/r2c11_bench_16k.err; echo "exit $?" >> gpurun_out/r2c11_bench_16k.err )
tail -2 gpurun_out/r2c11_bench_16k.err; cat gpurun_out/r2c11_bench_16k.json | cut -c1-1200
