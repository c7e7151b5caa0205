#!/bin/bash
# A/B session for the one-row-tile-per-CTA attention kernel (LONGSPEC_TC1 = 0 trios, 1 one row tile per CTA, 2 clusters of two): parity per mode, per-launch timings, cycle traces, bench (profiles/r02b_attn_tc1_*).
# Every step runs under its own timeout; a hang in a parity run aborts the session (exit 124) before anything longer starts.
set -u
mkdir -p gpurun_out
for m in 1 2; do
  LONGSPEC_TC1=$m timeout 150 python -m pytest tests/test_attention_tc_gpu.py -x -q > gpurun_out/tc1_parity_mode$m.log 2>&1
  rc=$?; echo "mode $m parity exit $rc" | tee -a gpurun_out/tc1_parity_mode$m.log; tail -4 gpurun_out/tc1_parity_mode$m.log | cut -c1-400
  if [ $rc -eq 124 ]; then echo "HANG in mode $m"; [ $m -eq 1 ] && exit 124; fi
done
for m in 0 1 2; do
  for n in 131072 16384; do
    ( LONGSPEC_TC1=$m timeout 120 python tools/prof_attn.py $n 12 > gpurun_out/tc1_prof_mode${m}_$n.txt 2>&1 ); echo "mode $m N=$n: $(tail -1 gpurun_out/tc1_prof_mode${m}_$n.txt | cut -c1-200)"; grep "ls_attn_partial alone" gpurun_out/tc1_prof_mode${m}_$n.txt
  done
done
for m in 1 2; do
  for bx in 0 12; do
    ( LONGSPEC_TC1=$m BX=$bx timeout 120 python tools/trace_attn.py 131072 > gpurun_out/tc1_trace_mode${m}_bx$bx.txt 2>&1 ); echo "--- trace mode $m bx $bx"; tail -16 gpurun_out/tc1_trace_mode${m}_bx$bx.txt | cut -c1-200
  done
done
( LONGSPEC_TC1=2 timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_seams_gpu.py -x -q > gpurun_out/tc1_e2e_mode2.log 2>&1; echo "e2e exit $?" >> gpurun_out/tc1_e2e_mode2.log ); tail -3 gpurun_out/tc1_e2e_mode2.log | cut -c1-300
for m in 0 2; do
  ( LONGSPEC_TC1=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/tc1_bench_mode$m.json 2> gpurun_out/tc1_bench_mode$m.err ); cut -c1-400 gpurun_out/tc1_bench_mode$m.json
done
