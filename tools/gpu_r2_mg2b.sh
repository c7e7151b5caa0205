#!/bin/bash
# vocab-sharded lm_head: kernel-level emulation tests (1 GPU) + 2-GPU token identity + bench N=2
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_tree_ops_gpu.py tests/test_seams_gpu.py tests/test_driver_gpu.py -x -q 2>&1 | tail -12 ) | tee gpurun_out/r2mg2b_pytest.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
echo "=== check_multigpu tiny-g4 (peer + vocab shard)"
( run tools/check_multigpu.py tiny-g4 3000 ) 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/r2mg2b_check.log
echo "=== bench N=2"
( run bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2mg2b_bench.json 2> gpurun_out/r2mg2b_bench.err ); tail -2 gpurun_out/r2mg2b_bench.err; cut -c1-400 gpurun_out/r2mg2b_bench.json
echo "=== bench N=2 no vocab shard"
( LONGSPEC_VOCAB_SHARD=0 run bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2mg2b_bench_novs.json 2> gpurun_out/r2mg2b_bench_novs.err ); tail -2 gpurun_out/r2mg2b_bench_novs.err; cut -c1-400 gpurun_out/r2mg2b_bench_novs.json
