set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_long.py -m gpu -x -q > gpurun_out/r2_tests20.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests20.log
tail -5 gpurun_out/r2_tests20.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b n128 --length 120 --series 12000000 --max-lag 8 --steps 10 --warmup 3 --no-cpu --no-e2e
b n256 --length 240 --series 6000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n512 --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024 --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
timeout 300 python tools/long_probe.py > gpurun_out/long_probe.log 2>&1; tail -5 gpurun_out/long_probe.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_sub --launch-skip 3 -c 1 -o gpurun_out/prof_sub3_r02 -f python bench.py --length 480 --series 3000000 --max-lag 15 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_sub3.log 2>&1
