set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log
tail -4 gpurun_out/r2_tests3.log
timeout 600 python bench.py --workload c4 --ungrouped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4u_b.json 2> gpurun_out/r2_c4u_b.err
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c4g_b.json 2> gpurun_out/r2_c4g_b.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_big -c 1 -o gpurun_out/prof_big_v2 python bench.py --workload c4 --ungrouped --series 296000 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_big_v2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_c4*_b.json')):
    j=json.loads(open(f).read().strip().splitlines()[-1]); r=j['roofline']; c=j['config']
    print(f, 'step %.2f kernel %.2f frac %.3f refined %.0f rescored %.0f tail %.2f' % (j['ms_per_step'], r['kernel_ms'], r['frac'], c['refined_per_step'], c['rescored_per_step'], c['tail_ms']))
PY
