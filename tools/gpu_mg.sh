#!/bin/bash
# multi-GPU session (gpurun --gpus N): sharded tokens == single-GPU tokens (8/2-head toy model = trios), then bench.py --gpus N
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_mg.sh 2'
set -u
mkdir -p gpurun_out
N=${1:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
( run tools/check_multigpu.py tiny-g4 3000 ) 2>&1 | grep -E "shard|check ok|rror|ssert" | tail -$((N + 2)) | tee gpurun_out/mg${N}_check.log
( run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench.out 2> gpurun_out/mg${N}_bench.err ); grep '^{' gpurun_out/mg${N}_bench.out > gpurun_out/mg${N}_bench.json
tail -2 gpurun_out/mg${N}_bench.err | cut -c1-200; python - <<PY
import json
j=json.load(open("gpurun_out/mg${N}_bench.json")); r=j["roofline"]
print("N=$N ms/round %.3f e2e ms %.3f tok/s %.1f attn us %.1f frac %.3f | %s | ranks_agree %s checked %s match %s clocks %s" % (j["ms_per_step"], j["e2e"]["ms_per_step"], j["value"], r["us_per_launch"], r["frac"], j["config"]["exchange"], j["ranks_agree"], j["tokens_checked"], j["tokens_match_greedy"], j["clocks"]["sm_mhz"]))
PY
