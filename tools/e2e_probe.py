"""Where the end-to-end step's time goes: append (H2D) / NewBatch / Run, a few repetitions each."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import numpy as np, torch
import muse_b200 as mb
S, N, SEED = 1_000_000, 1440, 20261018
ctx = mb.Context(0)
store = mb.DeviceStore(ctx, N, 2, S)
store.append_synthetic(S, SEED, 0)
ref = mb.synth_reference(SEED, N)
host = torch.empty((S, N), dtype=torch.float64, pin_memory=True)
store.read_rows_ptr(0, S, host.data_ptr())
ids = torch.zeros((S, 2), dtype=torch.int32, pin_memory=True)
st2 = mb.DeviceStore(ctx, N, 2, S)
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for it in range(4):
    a, _ = T(lambda: st2.clear())
    b, _ = T(lambda: st2.append_host_ptr(host.data_ptr(), S, N, ids.data_ptr()))
    c, b2 = T(lambda: mb.DeviceBatch(ctx, st2, ref))
    d, r = T(lambda: b2.run([], 60, 100, 0.5))
    e, _ = T(lambda: b2.close())
    print("iter %d: clear %.2f  append %.2f (%.1f GB/s)  NewBatch %.2f  Run %.2f  close %.2f ms" % (it, a, b, S * N * 8 / b / 1e6, c, d, e), flush=True)
# plain torch copy of the same pinned buffer for comparison
d = torch.empty((S, N), dtype=torch.float64, device="cuda")
for it in range(3):
    t, _ = T(lambda: d.copy_(host, non_blocking=True))
    print("torch copy_ of the same pinned rows: %.2f ms (%.1f GB/s)" % (t, S * N * 8 / t / 1e6), flush=True)
# ids handling on the host side of append
t0 = time.perf_counter(); x = ids.numpy().copy(); print("host copy of ids %.2f ms" % ((time.perf_counter() - t0) * 1e3))
