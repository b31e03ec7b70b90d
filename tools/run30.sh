set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests30.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests30.log
tail -4 gpurun_out/r2_tests30.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke30.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke30.log
tail -2 gpurun_out/r2_smoke30.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 200 gpurun_out/r2_n1_$name.json; echo; tail -2 gpurun_out/r2_n1_$name.err; }
b c4g --workload c4 --steps 5 --warmup 3
b c4u --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c3 --steps 20 --warmup 3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_ungrouped_r02 -f python bench.py --workload c4 --ungrouped --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big --launch-skip 3 -c 1 -o gpurun_out/prof_big_grouped_r02 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bigg.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4g_launches.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list4.log 2>&1
