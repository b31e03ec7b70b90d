set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests22.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests22.log
tail -5 gpurun_out/r2_tests22.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b n512_f --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024_f --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
export MUSE_B200_LIB=build/variants/lib_bigpair.so
b c4u_pair --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c4g_pair --workload c4 --steps 5 --warmup 3 --no-cpu
unset MUSE_B200_LIB
b c4u_base --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
