set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2_tests24.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests24.log
tail -5 gpurun_out/r2_tests24.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c3_w --steps 20 --warmup 3 --no-cpu --no-e2e
b c4u_w --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c4g_w --workload c4 --steps 5 --warmup 3 --no-cpu
b c5_w --workload c5 --steps 5 --warmup 3 --no-cpu
b n512_w --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_top10k_w --top-n 10000 --steps 10 --warmup 3 --no-cpu --no-e2e
