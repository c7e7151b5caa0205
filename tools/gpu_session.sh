#!/bin/bash
# One GPU-box session (1 GPU) producing the evidence kept under profiles/:
#   parity suite, smoke, bench (ours + reference arm), FA2 baseline of the reference's attention, per-kernel round
#   breakdown, ncu launch list of the bench command, ncu --set full of the dominant kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'      then: python tools/summarize_profiles.py r02
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log )
tail -2 gpurun_out/smoke.log | cut -c1-300
( timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err )
cut -c1-600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err )
cut -c1-300 gpurun_out/bench_reference_arm.json
( timeout 300 python tools/ref_gpu_attn.py 131072 12 > gpurun_out/ref_gpu_attn_128k.log 2>&1 ); tail -1 gpurun_out/ref_gpu_attn_128k.log | cut -c1-400
( timeout 300 python tools/ref_gpu_attn.py 16384 30 > gpurun_out/ref_gpu_attn_16k.log 2>&1 )
( timeout 300 python tools/prof_attn.py 131072 12 > gpurun_out/prof_attn_128k.txt 2>&1 ); tail -3 gpurun_out/prof_attn_128k.txt
( timeout 300 python tools/prof_attn.py 16384 30 > gpurun_out/prof_attn_16k.txt 2>&1 ); tail -3 gpurun_out/prof_attn_16k.txt
( timeout 600 python tools/prof_round.py 131072 > gpurun_out/prof_round_128k.txt 2>&1 )
# launch list of the TIMED rounds only (bench.py brackets them with cudaProfilerStart/Stop: the prefill alone is > 10^5 launches)
( LONGSPEC_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1 )
wc -l gpurun_out/launches_bench.csv
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 2 -o gpurun_out/attn_tc_full -f \
    python tools/prof_attn.py 131072 3 > gpurun_out/ncu_full.log 2>&1 )
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grow_phase|attn_partial_kernel|tree_accept|rope_append" -s 12 -c 8 -o gpurun_out/draft_kernels_full -f \
    python tools/prof_round.py 16384 > gpurun_out/ncu_draft.log 2>&1 )
ls -la gpurun_out | head -40
