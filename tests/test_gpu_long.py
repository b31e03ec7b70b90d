"""Series whose FFT length is above 16384 (kernels_long.cu): go-muse has no length limit (muse_batch.go:23-52), and its
own BenchmarkXCorrWithX runs 16385 samples at n = 32768 (xcorr_test.go:330-348).  Needs a B200."""
import numpy as np
import pytest

import muse_b200 as mb
from oracle import c_oracle as co
from oracle import muse_oracle as mo

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    return mb.default_context(0)


def _rows(rng, S, N):
    Y = rng.random((S, N)) - 0.5
    for i in range(0, S, 2):
        m = int(rng.integers(N // 4, 3 * N // 4))
        w = int(rng.integers(5, N // 16))
        Y[i, m:m + w] += rng.uniform(2.0, 30.0) * (1 if i % 4 else -1)
    return Y


@pytest.mark.parametrize("N,S", [(16385, 37), (20000, 8), (32768, 5), (40001, 3), (70000, 2)])
def test_score_all_matches_oracle_above_the_fused_lengths(ctx, N, S):
    rng = np.random.default_rng(N + S)
    Y = _rows(rng, S, N)
    Y[S - 1] = 4.25                       # std == 0 -> (0, 0), xcorr.go:165-168
    ref = np.zeros(N)
    ref[N // 2:N // 2 + N // 20] = 3.0
    ref += 0.2 * (rng.random(N) - 0.5)
    store = mb.DeviceStore(ctx, N, 0, S)
    store.append(Y)
    b = mb.DeviceBatch(ctx, store, ref)
    assert b.fft_len() == mo.next_pow_of2(N) > 16384
    for signed in (False, True):
        sc, lg = b.score_all(signed_scores=signed)
        wsc, wlg = co.score_all(ref, Y, signed=signed)
        assert np.max(np.abs(sc - wsc)) <= SCORE_TOL
        np.testing.assert_array_equal(lg, wlg)
        assert sc[S - 1] == 0.0 and lg[S - 1] == 0
    # the whole cross-correlation of one series (xCorrWithX, xcorr.go:160-197)
    cc, std_zero = b.xcorr(0)
    rc, want_cc, want_lag, want_mv = co.xcorr_with_x(ref, Y[0])
    assert rc == 0 and not std_zero and cc.shape == want_cc.shape
    assert np.max(np.abs(cc - want_cc)) <= SCORE_TOL
    cc, std_zero = b.xcorr(S - 1)
    assert std_zero and cc is None


def test_batch_run_above_the_fused_lengths(ctx):
    # Batch.Run (muse_batch.go:99-130) ungrouped and grouped, Muse.Run's signed scores, at n = 32768
    rng = np.random.default_rng(7)
    N, S = 16385, 61
    Y = _rows(rng, S, N)
    ref = np.zeros(N)
    ref[8000:8400] = 2.0
    ref += 0.1 * (rng.random(N) - 0.5)
    ids = np.stack([np.arange(S) // 4, np.arange(S) % 4], axis=1).astype(np.int32)
    store = mb.DeviceStore(ctx, N, 2, S)
    store.append(Y, ids)
    b = mb.DeviceBatch(ctx, store, ref)
    for keys, gids in (([], None), ([0], ids[:, 0].astype(np.int64)), ([1], ids[:, 1].astype(np.int64))):
        for thr, max_lag in ((0.0, 20000), (0.02, 4000)):
            sc, lg, ix = b.run(keys, max_lag, 10, thr)
            wsc, wlg, wix = co.batch_run(ref, Y, gids, max_lag, 10, thr)
            assert len(sc) == len(wsc)
            assert np.max(np.abs(sc - wsc), initial=0.0) <= SCORE_TOL
            np.testing.assert_array_equal(lg, wlg)
            np.testing.assert_array_equal(ix, wix)
    assert b.timing().mode == mb.MODE_EXACT
    # a reference whose std is zero is "Invalid input query" here too (muse_batch.go:38-41)
    with pytest.raises(mb.MuseError) as e:
        mb.DeviceBatch(ctx, store, np.full(N, 2.0))
    assert e.value.code == mb.MUSE_ERR_STDDEV_ZERO


def test_facade_run_with_long_series(ctx):
    # NewSeries / NewGroup / NewBatch / Run / Fetch with 20 000 samples per series against the Python oracle's classes
    rng = np.random.default_rng(11)
    N, S = 20000, 12
    Y = _rows(rng, S, N)
    ref = np.zeros(N)
    ref[9000:9600] = 1.0
    ref += 0.05 * (rng.random(N) - 0.5)
    g = mb.NewGroup("long")
    g.Add(*[mb.NewSeries(Y[i], mb.NewLabels({"graph": "g%d" % (i // 3), "host": "h%d" % (i % 3)})) for i in range(S)])
    b = mb.NewBatch(mb.NewSeries(ref), g, mb.NewResults(25000, 5, 0.0, mb.SignFilter_ANY), 2)
    assert b.Run(["graph"]) is None
    got, _ = b.Results.Fetch()
    wsc, wlg, wix = mo.batch_run_arrays(ref, Y, np.arange(S) // 3, 25000, 5, 0.0)
    assert len(got) == len(wsc) == 4
    for s, ws, wl, wi in zip(got, wsc, wlg, wix):
        assert abs(s.PercentScore - ws) <= SCORE_TOL and s.Lag == wl
        assert s.Labels is g.series[int(wi)].Labels()


@pytest.mark.parametrize("lx,ly,n", [(4097, 4000, 0), (10007, 10007, 10007), (16385, 16385, 32768), (50000, 49999, 65536),
                                     (100003, 100003, 0), (300000, 300000, 300000)])
@pytest.mark.parametrize("normalize", [False, True])
def test_generic_xcorr_by_fft_matches_oracle(ctx, lx, ly, n, normalize):
    # xCorr (xcorr.go:102-153) above 4096 lags: FFT passes; lengths that are not powers of two fold the linear correlation
    rng = np.random.default_rng(lx + ly + n)
    x, y = rng.random(lx), rng.random(ly)
    y[ly // 3] += 25.0
    x[lx // 2] += 25.0
    cc, lag, mv = mb.xCorr(x, y, n, normalize, ctx)
    want_cc, want_lag, want_mv = mo.x_corr(x, y, n, normalize)
    assert cc.shape == want_cc.shape
    scale = max(1.0, float(np.max(np.abs(want_cc))))
    assert np.max(np.abs(cc - want_cc)) <= SCORE_TOL * scale
    assert lag == want_lag and abs(mv - want_mv) <= SCORE_TOL * scale
