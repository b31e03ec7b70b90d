set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py "tests/test_gpu_screen.py::test_multi_query_randomised_against_separate_runs" -m gpu -x -q > gpurun_out/r2_tc4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tc4.log
tail -3 gpurun_out/r2_tc4.log
for v in main rw6 rw10; do
  if [ $v = main ]; then L=""; else L=$PWD/build/variants/lib_$v.so; fi
  MUSE_B200_LIB=$L timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n1_$v.json 2> gpurun_out/r2_c5_n1_$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c5_launches_c.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5_c.log 2>&1
