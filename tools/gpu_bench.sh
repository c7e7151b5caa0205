#!/bin/bash
# GPU session: parity tests, smoke, bench, launch list of the bench command, ncu --set full of the dominant kernel.
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log )
tail -2 gpurun_out/smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_FLAGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err )
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
( timeout 600 python tools/prof_round.py 131072 > gpurun_out/prof_round_128k.txt 2>&1 )
head -36 gpurun_out/prof_round_128k.txt
if [ "${NCU:-1}" = "1" ]; then
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1 )
wc -l gpurun_out/launches_bench.csv
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 2 -o gpurun_out/attn_tc_full -f \
    python tools/prof_attn.py 131072 3 > gpurun_out/ncu_full.log 2>&1 )
fi
ls -la gpurun_out | head -30
