cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for g in "" "1"; do
for w in "--ungrouped" ""; do
  if [ -n "$g" ]; then export MUSE_B200_LIB=$PWD/build/variants/lib_smemtw.so; else unset MUSE_B200_LIB; fi
  python bench.py --workload c4 $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('generic=$g $w', 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'step %.3f' % j['ms_per_step'])
" | tee -a gpurun_out/ab29.log
done; done; done
