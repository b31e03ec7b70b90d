set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_long.py -m gpu -x -q > gpurun_out/r2_tests16_long.log 2>&1; echo "long rc=$?" >> gpurun_out/r2_tests16_long.log
tail -15 gpurun_out/r2_tests16_long.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke16.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke16.log
tail -3 gpurun_out/r2_smoke16.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests16.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests16.log
tail -6 gpurun_out/r2_tests16.log
timeout 600 python bench.py > gpurun_out/r2_n1_default16.json 2> gpurun_out/r2_n1_default16.err; tail -c 300 gpurun_out/r2_n1_default16.json
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:score_exact_kernel<10, 4, 0" --launch-skip 3 -c 1 -o gpurun_out/prof_exact_r02 -f python bench.py --mode exact --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_exact.log 2>&1
tail -3 gpurun_out/ncu_exact.log
timeout 300 python tools/long_probe.py > gpurun_out/long_probe.log 2>&1; tail -8 gpurun_out/long_probe.log
