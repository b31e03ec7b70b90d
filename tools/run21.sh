set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2_tests21.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests21.log
tail -5 gpurun_out/r2_tests21.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 300 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c3_pair --steps 20 --warmup 3 --no-cpu --no-e2e
b n128_pair --length 120 --series 12000000 --max-lag 8 --steps 10 --warmup 3 --no-cpu --no-e2e
b n256_pair --length 240 --series 6000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n512_pair --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024_pair --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_1441_pair --length 1441 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_top10k_pair --top-n 10000 --steps 10 --warmup 3 --no-cpu --no-e2e
timeout 600 python tools/screen_error_survey.py > gpurun_out/survey.log 2>&1; tail -3 gpurun_out/survey.log
