set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests2.log
tail -15 gpurun_out/r2_tests2.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big -c 1 -o gpurun_out/prof_big_v1 python bench.py --workload c4 --ungrouped --series 296000 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_big_v1.log 2>&1
tail -3 gpurun_out/ncu_big_v1.log
ls -la gpurun_out/*.ncu-rep
