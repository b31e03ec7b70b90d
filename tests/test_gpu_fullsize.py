"""BASELINE.json configurations at FULL size through the C ABI: properties that do not need an
oracle pass over the whole store (the oracle checks the returned rows and a random sample), plus
the C2 shape (BenchmarkMuseBatchRunLarge, muse_batch_test.go:137-190) against the oracle.  Needs a B200."""
import time

import numpy as np
import pytest

import muse_b200 as mb
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

SEED = 20261018


@pytest.fixture(scope="module")
def ctx():
    return mb.default_context(0)


@pytest.fixture(scope="module")
def c3(ctx):
    S, N = 1_000_000, 1440
    store = mb.DeviceStore(ctx, N, 2, S)
    store.append_synthetic(S, SEED, 0)
    ref = mb.synth_reference(SEED, N)
    return store, ref, mb.DeviceBatch(ctx, store, ref)


def test_c3_screened_run_is_the_exact_run(c3):
    # 1 M series x 1440, maxLag 60, topN 100, threshold 0.5 (and 0.0): bit-identical to scoring everything in fp64
    store, ref, b = c3
    for thr in (0.5, 0.0):
        e = b.run([], 60, 100, thr, mode=mb.MODE_EXACT)
        s = b.run([], 60, 100, thr, mode=mb.MODE_AUTO)
        t = b.timing()
        assert t.mode == mb.MODE_SCREEN and t.n_rescored < 20_000 and t.n_refined < 100_000
        for x, y in zip(e, s):
            np.testing.assert_array_equal(x, y)
        sc, lg, ix = s
        assert len(sc) == 100
        assert np.all(np.diff(sc) <= 0)                       # results.go:81-85: descending
        assert np.all(np.abs(lg) <= 60) and np.all(sc >= thr) and np.all(sc <= 1.0)
        assert len(set(ix.tolist())) == 100
        # idempotent: a second run returns the same thing (no state leaks between runs)
        for x, y in zip(s, b.run([], 60, 100, thr, mode=mb.MODE_AUTO)):
            np.testing.assert_array_equal(x, y)


def test_c3_result_rows_against_the_oracle(c3):
    # the oracle scores the 100 returned rows and 3000 random others: returned scores/lags agree, and
    # no sampled row that passes the filter beats the cut-off
    store, ref, b = c3
    sc, lg, ix = b.run([], 60, 100, 0.5)
    rows = np.stack([store.read_row(int(i)) for i in ix])
    wsc, wlg = co.score_all(ref, rows)
    assert np.max(np.abs(sc - wsc)) <= 1e-9
    np.testing.assert_array_equal(lg, wlg)
    rng = np.random.default_rng(5)
    sample = np.setdiff1d(rng.integers(0, store.size(), 3000), ix)
    rows = np.stack([store.read_row(int(i)) for i in sample])
    ssc, slg = co.score_all(ref, rows)
    passing = (np.abs(slg) <= 60) & (ssc >= 0.5)
    assert np.all(ssc[passing] <= sc[-1] + 1e-12)
    # top_n nests: the top 10 is the head of the top 100
    s10 = b.run([], 60, 10, 0.5)
    np.testing.assert_array_equal(s10[2], ix[:10])
    np.testing.assert_array_equal(s10[0], sc[:10])


def test_c3_sharded_merge_equals_whole(ctx, c3):
    # two half-size shards with global offsets -> partials -> merge == the 1 M store (multi-GPU data path on one GPU)
    store, ref, b = c3
    want = b.run([], 60, 100, 0.5)
    half = store.size() // 2
    parts = []
    for k in range(2):
        st = mb.DeviceStore(ctx, 1440, 2, half)
        st.append_synthetic(half, SEED, k * half)
        st.set_global_offset(k * half)
        parts.append(mb.DeviceBatch(ctx, st, ref).run_partial([], 60, 100, 0.5))
        del st
    got = mb.merge_partials(np.concatenate(parts), 60, 100, 0.5)
    for x, y in zip(want, got):
        np.testing.assert_array_equal(x, y)


def test_c4_shape_screened_run_is_the_exact_run(ctx):
    # C4's per-series shape (one week at one sample per minute: N = 10080, n = 16384, maxLag 240, topN 100) on
    # 100 k synthetic series (8 GB): the block screening kernel + fused second stage against scoring everything in fp64
    S, N = 100_000, 10080
    store = mb.DeviceStore(ctx, N, 2, S)
    store.append_synthetic(S, SEED, 0)
    ref = mb.synth_reference(SEED, N)
    b = mb.DeviceBatch(ctx, store, ref)
    for thr in (0.5, 0.0):
        e = b.run([], 240, 100, thr, mode=mb.MODE_EXACT)
        te = b.timing()
        s = b.run([], 240, 100, thr, mode=mb.MODE_AUTO)
        t = b.timing()
        assert t.mode == mb.MODE_SCREEN and t.n_rescored < 0.1 * S
        for x, y in zip(e, s):
            np.testing.assert_array_equal(x, y)
        assert len(s[0]) == 100 and np.all(np.abs(s[1]) <= 240)
    print("C4 shape, %d series: exact %.2f ms, screened %.2f ms (%d refined, %d exact)"
          % (S, te.total_ms, t.total_ms, t.n_refined, t.n_rescored))
    rows = np.stack([store.read_row(int(i)) for i in s[2][:20]])
    wsc, wlg = co.score_all(ref, rows)
    assert np.max(np.abs(s[0][:20] - wsc)) <= 1e-9
    np.testing.assert_array_equal(s[1][:20], wlg)


def test_c4_grouped_by_two_labels_is_the_exact_run_and_the_oracle(ctx):
    # BASELINE.json configs[3] as benched: N = 10080, grouped by [graph, host] (100 series per group), maxLag 240, topN 100,
    # on 100 k synthetic series: the screened run (running per-group lower bound, muse_batch.go:87-89) must be the all-exact
    # run bit for bit, and both must be the oracle's answer (group.go:76-104 + muse_batch.go:56-93 + results.go:46-87)
    S, N = 100_000, 10080
    store = mb.DeviceStore(ctx, N, 3, S)
    store.append_synthetic(S, SEED, 0)
    store.set_synthetic_labels([10000, 100, 1], [1000, 100, 100])
    ref = mb.synth_reference(SEED, N)
    b = mb.DeviceBatch(ctx, store, ref)
    e = b.run([0, 1], 240, 100, 0.5, mode=mb.MODE_EXACT)
    s = b.run([0, 1], 240, 100, 0.5, mode=mb.MODE_AUTO)
    t = b.timing()
    assert t.mode == mb.MODE_SCREEN and t.n_rescored < 0.1 * S and t.n_refined < 0.6 * S
    for x, y in zip(e, s):
        np.testing.assert_array_equal(x, y)
    print("C4 grouped, %d series: screened %.2f ms (%d refined, %d exact)" % (S, t.total_ms, t.n_refined, t.n_rescored))
    Y = store.read_rows(0, S)
    wsc, wlg, wix = co.batch_run(ref, Y, (np.arange(S) // 100).astype(np.int64), 240, 100, 0.5)
    assert len(wsc) == len(s[0]) == 100
    assert np.max(np.abs(s[0] - wsc)) <= 1e-9
    np.testing.assert_array_equal(s[1], wlg)
    np.testing.assert_array_equal(s[2], wix)


def test_c2_benchmark_shape_matches_oracle(ctx):
    # muse_batch_test.go:137-190: 100 graphs x 50 hosts, N = 480, noise rows, Run(["graph"]), maxLag 10, topN 20, thr 0
    rng = np.random.default_rng(42)
    G, H, N = 100, 50, 480
    S = G * H
    Y = 0.1 * (rng.random((S, N)) - 0.5)
    ref = 0.1 * (rng.random(N) - 0.5)
    ids = np.stack([np.arange(S) // H, np.arange(S) % H], axis=1).astype(np.int32)
    store = mb.DeviceStore(ctx, N, 2, S)
    store.append(Y, ids)
    b = mb.DeviceBatch(ctx, store, ref)
    sc, lg, ix = b.run([0], 10, 20, 0.0)
    wsc, wlg, wix = co.batch_run(ref, Y, ids[:, 0].astype(np.int64), 10, 20, 0.0)
    assert len(sc) == len(wsc)
    assert np.max(np.abs(sc - wsc), initial=0.0) <= 1e-9
    np.testing.assert_array_equal(ix, wix)
    np.testing.assert_array_equal(lg, wlg)
    t0 = time.perf_counter()
    for _ in range(20):
        b.run([0], 10, 20, 0.0)
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print("C2 Batch.Run(['graph']): %.3f ms per run (README: 128 ms on 4 laptop cores)" % ms)
    assert ms < 50
