"""C5 probe (BASELINE.json configs[4] shape on ONE GPU): Q reference queries against the resident C3 store
through muse_multi_run.  Prints pair-samples/s.  usage: python tools/c5_probe.py [Q] [S]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import numpy as np
import muse_b200 as mb
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
N, SEED = 1440, 20261018
ctx = mb.Context(0)
store = mb.DeviceStore(ctx, N, 2, S)
store.append_synthetic(S, SEED, 0)
rng = np.random.default_rng(3)
refs = np.zeros((Q, N))
for q in range(Q):                                   # rect mids / widths varied
    mid, w = int(rng.integers(540, 900)), int(rng.integers(3, 21))
    refs[q, mid - w // 2: mid - w // 2 + w] = 1.5
    refs[q] += 0.1 * (rng.random(N) - 0.5)
mb.multi_run(store, refs, [], 60, 100, 0.5)          # warm-up: row statistics, scratch pool of every batch of a launch
ctx.synchronize()
REP = int(os.environ.get("C5_REP", "3"))
t0 = time.perf_counter()
for _ in range(REP):
    out = mb.multi_run(store, refs, [], 60, 100, 0.5)
dt = (time.perf_counter() - t0) / REP
print("C5 probe: %d refs x %d series x %d samples: %.1f ms (%.2f ms per query) = %.1f G pair-samples/s; results per query: %s"
      % (Q, S, N, dt * 1e3, dt * 1e3 / Q, Q * S * N / dt / 1e9, [len(o[0]) for o in out][:8]), flush=True)
if os.environ.get("C5_REFINED"):
    for q in range(min(Q, 8)):
        b = mb.DeviceBatch(ctx, store, refs[q])
        b.run([], 60, 100, 0.5)
        tm = b.timing()
        print("query %d alone: %.2f ms, refined %d, exact-scored %d" % (q, tm.total_ms, tm.n_refined, tm.n_rescored), flush=True)
        b.close()
