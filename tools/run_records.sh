# usage (inside gpurun, one GPU): bash tools/run_records.sh
# The single-GPU records, launch lists and ncu captures kept under profiles/ (copy from gpurun_out/ afterwards;
# python profiles/summarize.py gpurun_out/<capture>.ncu-rep profiles/<name>.txt <series per launch> makes the text summaries).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_n1_$name.json 2> gpurun_out/r2_n1_$name.err; tail -c 200 gpurun_out/r2_n1_$name.json; echo; }
b c3 --steps 20 --warmup 3
b c3_thr0 --threshold 0.0 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_top10k --top-n 10000 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_1441 --length 1441 --steps 10 --warmup 3 --no-cpu --no-e2e
b c3_rect --data rect --steps 5 --warmup 3 --no-cpu --no-e2e
b c4g --workload c4 --steps 5 --warmup 3
b c4u --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu
b c5 --workload c5 --steps 5 --warmup 3
b n128 --length 120 --series 12000000 --max-lag 8 --steps 10 --warmup 3 --no-cpu --no-e2e
b n256 --length 240 --series 6000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n512 --length 480 --series 3000000 --max-lag 15 --steps 10 --warmup 3 --no-cpu --no-e2e
b n1024 --length 1000 --series 1500000 --max-lag 30 --steps 10 --warmup 3 --no-cpu --no-e2e
b n4096 --length 2500 --series 600000 --max-lag 60 --steps 5 --warmup 3 --no-cpu --no-e2e
b n8192 --length 5000 --series 300000 --max-lag 120 --steps 5 --warmup 3 --no-cpu --no-e2e
timeout 300 python tools/long_probe.py > gpurun_out/long_probe.log 2>&1
timeout 600 python tools/screen_error_survey.py > gpurun_out/survey.log 2>&1
cap() { out=$1; kern=$2; shift 2; timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$kern" --launch-skip 3 -c 1 -o gpurun_out/$out -f python bench.py "$@" --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$out.log 2>&1; }
cap prof_warp_r02 score_screen_warp
cap prof_big_ungrouped_r02 score_screen_big --workload c4 --ungrouped
cap prof_big_grouped_r02 score_screen_big --workload c4
cap prof_sub3_r02 score_screen_sub --length 480 --series 3000000 --max-lag 15
cap prof_exact_r02 'score_exact_kernel<\(int\)10, \(int\)4, \(int\)0' --mode exact
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4g_launches.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list4.log 2>&1
