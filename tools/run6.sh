set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2_tc3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tc3.log
tail -3 gpurun_out/r2_tc3.log
for v in main rw10 rw8; do
  if [ $v = main ]; then L=""; else L=$PWD/build/variants/lib_$v.so; fi
  MUSE_B200_LIB=$L timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n1_$v.json 2> gpurun_out/r2_c5_n1_$v.err
done
MUSE_MULTI_TC=0 timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/r2_c5_n1_fp32.json 2> gpurun_out/r2_c5_n1_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c5_launches_b.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5_b.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:refine_multi -c 1 -o gpurun_out/prof_refine_v1 python bench.py --workload c5 --series 200000 --steps 1 --warmup 0 > gpurun_out/ncu_refine_v1.log 2>&1
