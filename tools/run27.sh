set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests27.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests27.log
tail -5 gpurun_out/r2_tests27.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c4u_r --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c4g_r --workload c4 --steps 5 --warmup 3 --no-cpu
b n4096_r --length 2500 --series 600000 --max-lag 60 --steps 5 --warmup 3 --no-cpu --no-e2e
b n8192_r --length 5000 --series 300000 --max-lag 120 --steps 5 --warmup 3 --no-cpu --no-e2e
