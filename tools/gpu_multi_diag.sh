#!/bin/bash
# Multi-GPU check with hard per-stage timeouts (run under gpurun --gpus N): sharded tokens == single-GPU tokens,
# then a short sharded bench.
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== sharded generation vs single GPU (CUDA graph)"; timeout -k 5 ${T1:-50} $TR --master-port 29503 tools/check_multigpu.py > gpurun_out/mg_graph.log 2>&1; echo "rc=$?"; grep -E "rank|check ok|rror|ssert" gpurun_out/mg_graph.log | tail -6
echo "== bench --gpus $N"; timeout -k 5 ${T2:-80} $TR --master-port 29504 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; cat gpurun_out/bench_${N}gpu.json | cut -c1-700; grep -E "rank|rror" gpurun_out/bench_${N}gpu.err | tail -4
