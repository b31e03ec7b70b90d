set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests8.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests8.log
tail -15 gpurun_out/r2_tests8.log
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n1_batched.json 2> gpurun_out/r2_c5_n1_batched.err
cut -c1-900 gpurun_out/r2_c5_n1_batched.json; tail -3 gpurun_out/r2_c5_n1_batched.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c5_launches_d.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5_d.log 2>&1
