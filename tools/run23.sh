set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests23.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests23.log
tail -5 gpurun_out/r2_tests23.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c4g --workload c4 --steps 5 --warmup 3
b c4u --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c3 --steps 20 --warmup 3
b c3_thr0 --threshold 0.0 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_top10k --top-n 10000 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_1441 --length 1441 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_rect --data rect --steps 5 --warmup 3 --no-cpu --no-e2e
b n128 --length 120 --series 12000000 --max-lag 8 --steps 10 --warmup 3 --no-cpu --no-e2e
b n256 --length 240 --series 6000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n512 --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024 --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_ungrouped_r02 -f python bench.py --workload c4 --ungrouped --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_grouped_r02 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigg.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_warp --launch-skip 3 -c 1 -o gpurun_out/prof_warp_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_warp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4g_launches.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list4.log 2>&1
