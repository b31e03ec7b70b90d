set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# the bench launches every kernel once on a tiny throw-away store and once for the cold run before the timed steps: skip those
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_warp --launch-skip 3 -c 1 -o gpurun_out/prof_warp_r02 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_warp.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_grouped_r02 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigg.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_ungrouped_r02 -f python bench.py --workload c4 --ungrouped --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:score_exact_kernel<10, 4, 0" --launch-skip 3 -c 1 -o gpurun_out/prof_exact_r02 -f python bench.py --mode exact --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_exact.log 2>&1
ls -la gpurun_out/prof_*_r02.ncu-rep
