cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests29.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests29.log
tail -3 gpurun_out/r2_tests29.log
for rep in 1 2; do
for g in "" "1"; do
for w in "--ungrouped" ""; do
  if [ -n "$g" ]; then export MUSE_BIG13_GENERIC=1; else unset MUSE_BIG13_GENERIC; fi
  python bench.py --workload c4 $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('generic=$g $w', 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'step %.3f' % j['ms_per_step'])
" | tee -a gpurun_out/ab29.log
done; done; done
