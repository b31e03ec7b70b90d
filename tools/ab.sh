#!/bin/bash
# run on the GPU box: bench each build/variants/lib_*.so twice, print kernel_ms / ms_per_step / refined
for rep in 1 2; do
for f in build/variants/lib_*.so; do
  MUSE_B200_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$f', 'kernel_ms %.4f' % j['roofline']['kernel_ms'], 'step %.4f' % j['ms_per_step'], 'refined', j['config'].get('refined_per_step'), 'rescored', j['config'].get('rescored_per_step'), 'tail', j['config'].get('tail_ms'))
"
done; done
