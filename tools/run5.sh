set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py "tests/test_gpu_screen.py::test_one_pass_multi_query_run_equals_exact_runs" "tests/test_gpu_screen.py::test_multi_query_randomised_against_separate_runs" tests/test_gpu_parity.py::test_multi_reference_run_equals_separate_batches -m gpu -x -q > gpurun_out/r2_tc2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tc2.log
tail -30 gpurun_out/r2_tc2.log
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n1_a.json 2> gpurun_out/r2_c5_n1_a.err
MUSE_MULTI_TC=0 timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/r2_c5_n1_fp32.json 2> gpurun_out/r2_c5_n1_fp32.err
cut -c1-700 gpurun_out/r2_c5_n1_a.json gpurun_out/r2_c5_n1_fp32.json; tail -3 gpurun_out/r2_c5_n1_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c5_launches_a.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5_a.log 2>&1
