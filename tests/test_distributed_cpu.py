"""World-size-2 gloo test of the multi-GPU host path on CPU: shard -> partial records ->
all-gather -> muse_merge_partials must equal the single-store answer.  The shard scores come
from the oracle here (no GPU in this container); the GPU suite checks the same equality with
partials produced by the CUDA path (tests/test_gpu_parity.py::test_sharded_partials_*)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    rng = np.random.default_rng(99)
    S, N = 600, 64
    ref = np.zeros(N)
    ref[28:36] = 1.5
    ref += 0.1 * (rng.random(N) - 0.5)
    Y = 0.1 * (rng.random((S, N)) - 0.5)
    for i in range(0, S, 2):
        m = int(rng.integers(10, 50))
        Y[i, m:m + int(rng.integers(2, 9))] += rng.uniform(0.5, 10)
    graph = rng.integers(0, 25, S)
    return ref, Y, graph


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "go-muse_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import muse_b200 as mb
    from oracle import muse_oracle as mo

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ref, Y, graph = _data()
    S = Y.shape[0]
    lo, hi = rank * S // world, (rank + 1) * S // world
    sc, lg = mo.score_series_batch(ref, Y[lo:hi])
    results = {}
    for name, grouped, max_lag, top_n, thr in (("u", False, 8, 10, 0.0), ("g", True, 8, 10, 0.2), ("g2", True, 32, 1000, 0.0)):
        if grouped:
            # every group representative of the shard, UNFILTERED
            best = {}
            for i in range(hi - lo):
                g = int(graph[lo + i])
                if g not in best or sc[i] > sc[best[g]]:
                    best[g] = i
            parts = np.zeros(len(best), dtype=mb.PARTIAL_DTYPE)
            for k, (g, i) in enumerate(sorted(best.items())):
                parts[k] = (g + 1, sc[i], lo + i, lg[i], 0)
            out = mb.allgather_merge(parts, max_lag, top_n, thr)
        else:
            # the shard's own filtered top_n
            ok = (np.abs(lg) <= max_lag) & (sc >= thr)
            idx = np.nonzero(ok)[0]
            idx = idx[np.argsort(-sc[idx], kind="stable")][:top_n]
            parts = np.zeros(len(idx), dtype=mb.PARTIAL_DTYPE)
            for k, i in enumerate(idx):
                parts[k] = (lo + i, sc[i], lo + i, lg[i], 0)
            out = mb.allgather_merge(parts, max_lag, top_n, thr, fixed_capacity=top_n)
        results[name] = [x.tolist() for x in out]
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_merge_matches_single_store():
    import muse_b200 as mb
    from oracle import muse_oracle as mo
    mb.build()
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, Y, graph = _data()
    first = {}
    gids = np.array([first.setdefault(int(g), len(first)) for g in graph])
    for name, g, max_lag, top_n, thr in (("u", None, 8, 10, 0.0), ("g", gids, 8, 10, 0.2), ("g2", gids, 32, 1000, 0.0)):
        wsc, wlg, wix = mo.batch_run_arrays(ref, Y, g, max_lag, top_n, thr)
        for rank in range(world):
            sc, lg, ix = (np.array(x) for x in got[rank][name])
            assert ix.tolist() == wix.tolist(), (name, rank)
            assert lg.tolist() == wlg.tolist()
            np.testing.assert_allclose(sc, wsc, rtol=0, atol=1e-12)
