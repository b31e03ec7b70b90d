"""A small run through every kernel family (for compute-sanitizer): warp kernel, block kernel, grouped
screening, device-side top-N, exact kernel, selection."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import numpy as np
import muse_b200 as mb
ctx = mb.Context(0)
rng = np.random.default_rng(1)
for N, S in ((1440, 3000), (1030, 700), (5000, 300), (10080, 120), (480, 2000)):
    Y = 0.1 * (rng.random((S, N)) - 0.5)
    for i in range(0, S, 3):
        m = int(rng.integers(N // 2 - N // 8, N // 2 + N // 8))
        Y[i, m:m + int(rng.integers(3, 20))] += rng.uniform(0.5, 40)
    ref = np.zeros(N); ref[N // 2 - 5:N // 2 + 5] = 1.5; ref += 0.1 * (rng.random(N) - 0.5)
    ids = np.stack([np.arange(S) // 7, np.arange(S) % 7], axis=1).astype(np.int32)
    st = mb.DeviceStore(ctx, N, 2, S); st.append(Y, ids)
    b = mb.DeviceBatch(ctx, st, ref)
    for mode in (mb.MODE_EXACT, mb.MODE_SCREEN):
        r1 = b.run([], 60, 50, 0.3, mode=mode)
        r2 = b.run([0], 60, 50, 0.3, mode=mode)
        p = b.run_partial([1], 60, 50, 0.3, mode=mode)
    if N > 1024:
        up, lo = b.screen_bounds(refine=True, max_lag=60)
    print("ok", N, S, len(r1[0]), len(r2[0]), len(p), flush=True)
    b.close(); st.close()
print("done")
