"""Multi-GPU step broken into phases (synchronised between phases, so this shows magnitudes, not overlap):
torchrun --nproc-per-node 2 tools/step_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import numpy as np, torch, torch.distributed as dist
import muse_b200 as mb
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, N, SEED = 1_000_000, 1440, 20261018
ctx = mb.Context(local)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
store = mb.DeviceStore(ctx, N, 2, S); store.append_synthetic(S, SEED, rank * S); store.set_global_offset(rank * S)
ref = mb.synth_reference(SEED, N)
b = mb.DeviceBatch(ctx, store, ref)
cap = 100
t = torch.empty(cap * 32, dtype=torch.uint8, device="cuda"); out = torch.empty(world * cap * 32, dtype=torch.uint8, device="cuda")
host = torch.empty(world * cap * 32, dtype=torch.uint8, pin_memory=True)
def now(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(8):
    dist.barrier(); t0 = now()
    b.run_partial_device(60, 100, 0.5, 0, mb.MODE_AUTO, t.data_ptr(), cap); c1 = time.perf_counter(); t1 = now()
    dist.all_gather_into_tensor(out, t); c2 = time.perf_counter(); t2 = now()
    host.copy_(out, non_blocking=True); torch.cuda.current_stream().synchronize(); t3 = time.perf_counter()
    r = mb.merge_partials(host.numpy().view(mb.PARTIAL_DTYPE), 60, 100, 0.5); t4 = time.perf_counter()
    tm = b.timing()
    if rank == 0 and it >= 3:
        print("run_partial_device: cpu %.0f us, gpu done %.0f us (lib total %.0f us: score %.0f tail %.0f) | all_gather cpu %.0f us, done %.0f us | d2h %.0f us | merge %.0f us"
              % ((c1 - t0) * 1e6, (t1 - t0) * 1e6, tm.total_ms * 1e3, tm.score_ms * 1e3, (tm.total_ms - tm.score_ms) * 1e3, (c2 - t1) * 1e6, (t2 - t1) * 1e6, (t3 - t2) * 1e6, (t4 - t3) * 1e6), flush=True)
# unsynchronised loop for comparison
dist.barrier(); t0 = now()
for _ in range(20):
    mb.allgather_merge_device(b, 60, 100, 0.5, 0)
t1 = now()
if rank == 0: print("allgather_merge_device loop: %.3f ms per step" % ((t1 - t0) / 20 * 1e3))
dist.barrier(); t0 = now()
for _ in range(20):
    b.run([], 60, 100, 0.5)
t1 = now()
if rank == 0: print("single-GPU run loop on rank 0 (other ranks run too): %.3f ms per step" % ((t1 - t0) / 20 * 1e3))
dist.destroy_process_group()
