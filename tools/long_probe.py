"""Timing probe of the long-series path (kernels_long.cu): S series of N samples through Batch.Run, and one generic xCorr."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import muse_b200 as mb

ctx = mb.default_context(0)
for N, S in ((16385, 20000), (100000, 4000), (1000000, 400)):
    store = mb.DeviceStore(ctx, N, 0, S)
    store.append_synthetic(S, 7, 0)
    ref = mb.synth_reference(7, N)
    b = mb.DeviceBatch(ctx, store, ref)
    b.run([], N // 40, 100, 0.5)
    t0 = time.perf_counter()
    for _ in range(3):
        sc, lg, ix = b.run([], N // 40, 100, 0.5)
    dt = (time.perf_counter() - t0) / 3
    print("N=%d S=%d n=%d: %.2f ms per Run = %.1f G series-samples/s, %d results, best %.6f" %
          (N, S, b.fft_len(), dt * 1e3, S * N / dt / 1e9, len(sc), sc[0] if len(sc) else 0.0), flush=True)
    del b, store
rng = np.random.default_rng(0)
x, y = rng.random(16385), rng.random(16385)
mb.xCorr(x, y, 32768, True, ctx)
t0 = time.perf_counter()
for _ in range(10):
    mb.xCorr(x, y, 32768, True, ctx)
print("xCorr 16385 samples at n = 32768 (BenchmarkXCorr's shape): %.3f ms per call" % ((time.perf_counter() - t0) / 10 * 1e3))
