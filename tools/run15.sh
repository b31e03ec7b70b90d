set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests15.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests15.log
tail -6 gpurun_out/r2_tests15.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 200 gpurun_out/r2_n1_$name.json; tail -2 gpurun_out/r2_n1_$name.err; }
b c4g_wide --workload c4 --steps 5 --warmup 3 --no-cpu
b c4u_wide --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
MUSE_BIG13=1 timeout 900 python bench.py --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_n1_c4u_big13.json 2> gpurun_out/r2_n1_c4u_big13.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_wide --launch-skip 3 -c 1 -o gpurun_out/prof_wide_ungrouped_r02 -f python bench.py --workload c4 --ungrouped --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_wideu.log 2>&1
