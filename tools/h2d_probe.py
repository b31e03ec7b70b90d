import os, time, torch, subprocess
print(subprocess.run("nvidia-smi topo -m | head -12; lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; nproc", shell=True, capture_output=True, text=True).stdout)
torch.cuda.set_device(0)
def bw(label):
    h = torch.empty(2 << 30, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty_like(h, device="cuda")
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = max(best, h.numel() / dt / 1e9)
    print(label, "H2D %.1f GB/s" % best, flush=True)
    del h, d
bw("default affinity %s" % (sorted(os.sched_getaffinity(0))[:4],))
allc = sorted(os.sched_getaffinity(0))
n = len(allc)
for part in range(4):
    cores = allc[part * n // 4:(part + 1) * n // 4]
    if not cores: continue
    os.sched_setaffinity(0, cores)
    bw("cores %d-%d" % (cores[0], cores[-1]))
os.sched_setaffinity(0, allc)
