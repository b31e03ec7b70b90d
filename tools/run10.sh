set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests10.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests10.log
tail -8 gpurun_out/r2_tests10.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; tail -2 gpurun_out/r2_n1_$name.err; }
b c3 --steps 20 --warmup 3
b c3_thr0 --steps 20 --warmup 3 --threshold 0.0 --no-e2e --no-cpu
b c3_top10k --steps 20 --warmup 3 --top-n 10000 --no-e2e --no-cpu
b c3_rect --steps 10 --warmup 3 --data rect --no-e2e --no-cpu
b c3_1441 --steps 20 --warmup 3 --length 1441 --no-e2e --no-cpu
b c4g --workload c4 --steps 5 --warmup 3
b c4u --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c5 --workload c5 --steps 5 --warmup 3
