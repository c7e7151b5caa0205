set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_elementwise_gpu.py tests/test_e2e_gpu.py tests/test_seams_gpu.py -x -q > gpurun_out/rope_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/rope_tests.log | cut -c1-300
( timeout 600 python tools/prof_round.py 16384 > gpurun_out/rope_prof_round_16k.txt 2>&1 ); grep -E "rope_append|add_rmsnorm|combine|silu|kernels:" gpurun_out/rope_prof_round_16k.txt | cut -c1-150
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/rope_bench.json 2> gpurun_out/rope_bench.err ); cut -c1-260 gpurun_out/rope_bench.json
