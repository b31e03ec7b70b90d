set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests32.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests32.log
tail -3 gpurun_out/r2_tests32.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 200 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c4g --workload c4 --steps 5 --warmup 3
b c4u --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b n4096 --length 2500 --series 600000 --max-lag 60 --steps 5 --warmup 3 --no-cpu --no-e2e
b n8192 --length 5000 --series 300000 --max-lag 120 --steps 5 --warmup 3 --no-cpu --no-e2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_ungrouped_r02 -f python bench.py --workload c4 --ungrouped --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_grouped_r02 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigg.log 2>&1
