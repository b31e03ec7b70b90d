set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests13.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests13.log
tail -8 gpurun_out/r2_tests13.log; grep "pageable append" gpurun_out/r2_tests13.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k pageable 2>&1 | grep -E "pageable|passed|failed"
timeout 600 python tools/screen_error_survey.py > gpurun_out/survey.log 2>&1; tail -3 gpurun_out/survey.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 200 gpurun_out/r2_n1_$name.json; tail -2 gpurun_out/r2_n1_$name.err; }
b c3_1441b --steps 20 --warmup 3 --length 1441 --no-e2e --no-cpu
b c4g_b --workload c4 --steps 5 --warmup 3 --no-cpu
b c4u_b --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
MUSE_B200_LIB=$PWD/build/variants/lib_minb3.so timeout 600 python bench.py --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_n1_c4u_minb3.json 2> gpurun_out/r2_n1_c4u_minb3.err
# ncu captures (one kernel each)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_warp -c 1 -o gpurun_out/prof_warp_r02 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_warp.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big -c 1 -o gpurun_out/prof_big_grouped_r02 python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_bigg.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big -c 1 -o gpurun_out/prof_big_ungrouped_r02 python bench.py --workload c4 --ungrouped --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_bigu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bounds_tc -c 1 -o gpurun_out/prof_bounds_tc_r02 python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:refine_multi -c 1 -o gpurun_out/prof_refine_r02 python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_refine.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_exact_kernel -c 1 -o gpurun_out/prof_exact_r02 python bench.py --mode exact --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_exact.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4g_launches.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
