set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sk in 0 16 32 48 64 80 96; do
for cfg in "240 6000000 15" "120 12000000 8" "480 3000000 15" "1000 1500000 30"; do
set -- $cfg
MUSE_SUB_SKEW=$sk timeout 300 python bench.py --length $1 --series $2 --max-lag $3 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('skew $sk N $1:', round(d['ms_per_step'],3), round(r['kernel_ms'],3), round(r['frac'],3))
" | tee -a gpurun_out/sub_skew.log
done
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_screen_sub --launch-skip 3 -c 1 -o gpurun_out/prof_sub1_r02 -f python bench.py --length 120 --series 12000000 --max-lag 8 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_sub1.log 2>&1
