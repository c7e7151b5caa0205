#!/bin/bash
# Final round-end GPU session (1 GPU): parity tests, bench (default flags), bench at the 16k config, ncu of the draft-side kernels.
set -u
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err )
cat gpurun_out/bench.json | cut -c1-400; tail -2 gpurun_out/bench.err
( LONGSPEC_BENCH_PROMPT=16384 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_16k.json 2> gpurun_out/bench_16k.err )
cat gpurun_out/bench_16k.json | cut -c1-300
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:"grow_phase|attn_partial_kernel|tree_accept|rope_append" -s 12 -c 8 -o gpurun_out/draft_kernels_full -f \
    python tools/prof_round.py 16384 > gpurun_out/ncu_draft.log 2>&1 )
tail -2 gpurun_out/ncu_draft.log
