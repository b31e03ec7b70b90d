set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_long.py -m gpu -x -q > gpurun_out/r2_tests17.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests17.log
tail -8 gpurun_out/r2_tests17.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 400 gpurun_out/r2_n1_$name.json | head -c 400; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b n512_sub --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024_sub --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
MUSE_BLOCK_SMALL=1 timeout 600 python bench.py --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_n1_n512_block.json 2> gpurun_out/r2_n1_n512_block.err; tail -c 300 gpurun_out/r2_n1_n512_block.json
MUSE_BLOCK_SMALL=1 timeout 600 python bench.py --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_n1_n1024_block.json 2> gpurun_out/r2_n1_n1024_block.err; tail -c 300 gpurun_out/r2_n1_n1024_block.json
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:score_exact_kernel<\(int\)10, \(int\)4, \(int\)0" --launch-skip 3 -c 1 -o gpurun_out/prof_exact_r02 -f python bench.py --mode exact --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_exact.log 2>&1
tail -3 gpurun_out/ncu_exact.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_sub --launch-skip 3 -c 1 -o gpurun_out/prof_sub3_r02 -f python bench.py --length 480 --series 3000000 --max-lag 15 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_sub3.log 2>&1
tail -3 gpurun_out/ncu_sub3.log
timeout 300 python tools/long_probe.py > gpurun_out/long_probe.log 2>&1; tail -5 gpurun_out/long_probe.log
MUSE_LONG_WORK_MB=64 timeout 300 python tools/long_probe.py > gpurun_out/long_probe64.log 2>&1; tail -5 gpurun_out/long_probe64.log
