set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests18.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests18.log
tail -8 gpurun_out/r2_tests18.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 600 gpurun_out/r2_n1_$name.json | head -c 600; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b n256_sub --length 240 --series 6000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n128_sub --length 120 --series 12000000 --max-lag 8 --steps 10 --warmup 3 --no-cpu --no-e2e
b n256_exact --length 240 --series 6000000 --max-lag 15 --steps 5 --warmup 3 --no-cpu --no-e2e --mode exact
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke18.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke18.log
tail -2 gpurun_out/r2_smoke18.log
