set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_c3_a.json 2> gpurun_out/r2_c3_a.err
timeout 600 python bench.py --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4u_a.json 2> gpurun_out/r2_c4u_a.err
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4g_a.json 2> gpurun_out/r2_c4g_a.err
MUSE_B200_LIB=$PWD/build/variants/lib_lb8.so timeout 600 python bench.py --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4u_lb8.json 2> gpurun_out/r2_c4u_lb8.err
MUSE_B200_LIB=$PWD/build/variants/lib_lb8.so timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4g_lb8.json 2> gpurun_out/r2_c4g_lb8.err
cat gpurun_out/r2_c3_a.json gpurun_out/r2_c4u_a.json gpurun_out/r2_c4g_a.json gpurun_out/r2_c4u_lb8.json gpurun_out/r2_c4g_lb8.json | cut -c1-1500
tail -3 gpurun_out/*.err
