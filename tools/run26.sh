set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/r2_tests26.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests26.log
tail -5 gpurun_out/r2_tests26.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c5_f --workload c5 --steps 5 --warmup 3
b c3_f --steps 20 --warmup 3 --no-cpu --no-e2e
