import os, sys
sys.path.insert(0, "/root/repo/go-muse_b200")
import numpy as np, muse_b200 as mb
S, N, SEED = 1_000_000, 1440, 20261018
ctx = mb.Context(0)
store = mb.DeviceStore(ctx, N, 2, S); store.append_synthetic(S, SEED, 0)
ref = mb.synth_reference(SEED, N)
b = mb.DeviceBatch(ctx, store, ref)
r = b.run([], 60, 100, 0.5); tm = b.timing()
print("rescored", tm.n_rescored, "refined", tm.n_refined, "100th", r[0][-1], "top", r[0][0])
sc, lg = b.score_all()
ok = (np.abs(lg) <= 60) & (sc >= 0.5)
s = np.sort(sc[ok])[::-1]
c = s[99]
for band in (0.0, 2e-4, 4.2e-4, 1e-3):
    print("band %g: in-window %d, any-lag %d" % (band, int((s >= c - band).sum()), int((sc >= c - band).sum())))
print("passing total", ok.sum(), "scores>=0.9:", int((s >= 0.9).sum()))
