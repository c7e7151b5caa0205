#!/bin/bash
# BASELINE.json configs 4 and 5 (and 2) through bench.py on one B200: golden record first (greedy stream), then the bench line.
set -u
mkdir -p gpurun_out
run() {   # name, env...
    name=$1; shift
    ( env "$@" timeout 600 python tools/make_bench_golden.py ) 2>&1 | tail -1 | cut -c1-300
    ( env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2cfg_${name}.json 2> gpurun_out/r2cfg_${name}.err; echo "exit $?" >> gpurun_out/r2cfg_${name}.err )
    tail -1 gpurun_out/r2cfg_${name}.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2cfg_${name}.json"))
    r=j["roofline"]
    print("${name}", "ms/round %.3f" % j["ms_per_step"], "tok/s %.1f" % j["value"], "e2e %.1f" % j["e2e"]["value"], "tau %.2f" % j["config"]["accepted_len_tau"],
          "attn us %.1f frac %.3f tflops %.0f" % (r["us_per_launch"], r["frac"], r["tensor_tflops"]), "checked", j["tokens_checked"], j["tokens_match_greedy"])
except Exception as e:
    print("${name} failed:", e)
PY
}
run qwq32b_4k   LONGSPEC_BENCH_MODEL=qwq-32b LONGSPEC_BENCH_PROMPT=4096
run qwq32b_24k  LONGSPEC_BENCH_MODEL=qwq-32b LONGSPEC_BENCH_PROMPT=24576
run vicuna13b_chain   LONGSPEC_BENCH_MODEL=vicuna-13b-16k LONGSPEC_BENCH_PROMPT=16384 "LONGSPEC_BENCH_TREE=1 1 1 1 1"
run vicuna13b_default LONGSPEC_BENCH_MODEL=vicuna-13b-16k LONGSPEC_BENCH_PROMPT=16384
run vicuna13b_wide    LONGSPEC_BENCH_MODEL=vicuna-13b-16k LONGSPEC_BENCH_PROMPT=16384 "LONGSPEC_BENCH_TREE=8 32 32 32 32"
