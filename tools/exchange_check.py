"""Multi-GPU check of the peer-memory exchange (torchrun --nproc-per-node N tools/exchange_check.py):
Exchange.run == NCCL all-gather path == host path, for several argument sets, then step timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import numpy as np, torch, torch.distributed as dist
import muse_b200 as mb
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, N, SEED = int(os.environ.get("CHECK_SERIES", "1000000")), 1440, 20261018
ctx = mb.Context(local)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
store = mb.DeviceStore(ctx, N, 2, S); store.append_synthetic(S, SEED, rank * S); store.set_global_offset(rank * S)
ref = mb.synth_reference(SEED, N)
b = mb.DeviceBatch(ctx, store, ref)
ex = mb.Exchange(ctx, 128)
ok = True
for max_lag, top_n, thr in ((60, 100, 0.5), (15, 10, 0.0), (60, 128, 0.9), (5, 1, 0.2), (60, 100, 0.5)):
    got = ex.run(b, max_lag, top_n, thr)
    want = mb.allgather_merge_device(b, max_lag, top_n, thr, 0)
    parts = b.run_partial([], max_lag, top_n, thr)
    host = mb.allgather_merge(parts, max_lag, top_n, thr, 0, fixed_capacity=top_n)
    same = got is not None and all(np.array_equal(x, y) for x, y in zip(got, want)) and all(np.array_equal(x, y) for x, y in zip(got, host))
    ok &= bool(same)
    if rank == 0:
        print("max_lag %d top_n %d thr %g: %s (n=%d, top score %.6f)" % (max_lag, top_n, thr, "identical" if same else "MISMATCH", len(got[0]) if got else -1, got[0][0] if got and len(got[0]) else float("nan")), flush=True)
# grouped: labels (i / 1000, i % 1000) of the GLOBAL index -> groups straddle the shards when grouped by "host"
gex = mb.Exchange(ctx, 4096)
for cols, max_lag, top_n, thr in (([1], 60, 100, 0.5), ([0], 60, 50, 0.0), ([1], 15, 1000, 0.2)):
    got = gex.run(b, max_lag, top_n, thr, key_cols=cols)
    parts = b.run_partial(cols, max_lag, top_n, thr)
    host = mb.allgather_merge(parts, max_lag, top_n, thr, 0)
    same = got is not None and all(np.array_equal(x, y) for x, y in zip(got, host))
    ok &= bool(same)
    if rank == 0:
        print("grouped by %s max_lag %d top_n %d thr %g: %s (n=%d)" % (cols, max_lag, top_n, thr, "identical" if same else "MISMATCH", len(got[0]) if got else -1), flush=True)
gex.close()
def now(): torch.cuda.synchronize(); return time.perf_counter()
for name, fn in (("peer-memory exchange", lambda: ex.run(b, 60, 100, 0.5)), ("NCCL all-gather from device memory", lambda: mb.allgather_merge_device(b, 60, 100, 0.5, 0))):
    for _ in range(3): fn()
    dist.barrier(); t0 = now()
    for _ in range(20): fn()
    t1 = now()
    t = torch.tensor([(t1 - t0) / 20 * 1e3], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%s: %.3f ms per step (max over ranks)" % (name, t.item()), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("ALL IDENTICAL" if flag.item() == 1 else "FAILED", flush=True)
ex.close(); b.close(); store.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
