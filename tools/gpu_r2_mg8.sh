#!/bin/bash
# N = 4 and N = 8 bench (peer exchange + vocab shard), token check inside bench.py (golden greedy stream + rank agreement)
set -u
mkdir -p gpurun_out
N=$1
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
( run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2mg${N}_bench.json 2> gpurun_out/r2mg${N}_bench.err ); tail -2 gpurun_out/r2mg${N}_bench.err | cut -c1-300; python - <<PY
import json
j=json.load(open("gpurun_out/r2mg${N}_bench.json")); r=j["roofline"]
print("N=$N ms/round %.3f e2e ms %.3f tok/s %.1f attn us %.1f frac %.3f exchange %s ranks_agree %s checked %s match %s" % (j["ms_per_step"], j["e2e"]["ms_per_step"], j["value"], r["us_per_launch"], r["frac"], j["config"]["exchange"], j["ranks_agree"], j["tokens_checked"], j["tokens_match_greedy"]))
PY
