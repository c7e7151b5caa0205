#!/bin/bash
# 2-GPU: peer exchange correctness (tokens == single GPU), NCCL fallback still correct, bench at N=2 both ways
set -u
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
echo "=== check_multigpu tiny-g4 (peer)"
( run tools/check_multigpu.py tiny-g4 3000 ) 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/r2mg2_check_peer.log
echo "=== check_multigpu tiny-g4 (nccl)"
( LONGSPEC_XCHG=nccl run tools/check_multigpu.py tiny-g4 3000 ) 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/r2mg2_check_nccl.log
echo "=== 2-GPU pytest"
( timeout 600 python -m pytest tests/test_e2e_gpu.py -q -k two_gpus 2>&1 | tail -3 ) | tee gpurun_out/r2mg2_pytest.log
echo "=== bench N=2 peer"
( run bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2mg2_bench_peer.json 2> gpurun_out/r2mg2_bench_peer.err ); tail -2 gpurun_out/r2mg2_bench_peer.err; cut -c1-700 gpurun_out/r2mg2_bench_peer.json
echo "=== bench N=2 nccl"
( LONGSPEC_XCHG=nccl run bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2mg2_bench_nccl.json 2> gpurun_out/r2mg2_bench_nccl.err ); tail -2 gpurun_out/r2mg2_bench_nccl.err; cut -c1-700 gpurun_out/r2mg2_bench_nccl.json
