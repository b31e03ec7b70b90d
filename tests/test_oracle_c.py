"""The C oracle (oracle/muse_oracle.c) against the numpy oracle and the reference KATs. CPU only."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import muse_oracle as mo


def test_c_kats(kats):
    for name in ("batch_run_simple", "batch_run_multi_dimensional"):
        k = kats[name]
        Y = np.array([e["y"] for e in k["comp"]], dtype=float)
        s, l = co.score_all(k["ref"]["y"], Y)
        s2, l2 = mo.score_series_batch(k["ref"]["y"], Y)
        np.testing.assert_allclose(s, s2, rtol=0, atol=1e-12)
        for i in range(len(l)):
            assert l[i] == l2[i] or name == "batch_run_simple" and i == 3 and l[i] in (-3, -2)
    with pytest.raises(ValueError):
        co.score_all([1.0, 1.0, 1.0, 1.0], np.ones((2, 4)))


@pytest.mark.parametrize("N", [2, 3, 8, 12, 31, 480, 1440])
def test_c_cc_vector_matches_numpy(N):
    rng = np.random.default_rng(N)
    ref = rng.standard_normal(N)
    y = rng.standard_normal(N) + 100.0
    rc, cc, lag, mv = co.xcorr_with_x(ref, y)
    X, n = mo.ref_spectrum(ref)
    cc2, lag2, mv2 = mo.x_corr_with_x(X, y, n)
    assert rc == 0
    np.testing.assert_allclose(cc, cc2, rtol=0, atol=1e-13)
    assert lag == lag2 and abs(mv - mv2) < 1e-13


def test_c_batch_run_matches_numpy():
    rng = np.random.default_rng(3)
    S, N = 3000, 480
    ref = np.zeros(N)
    ref[235:245] = 1.5
    ref += 0.1 * (rng.random(N) - 0.5)
    Y = 0.1 * (rng.random((S, N)) - 0.5)
    for i in range(0, S, 3):
        m = int(rng.integers(150, 330))
        w = int(rng.integers(3, 20))
        Y[i, m:m + w] += rng.uniform(0.5, 40)
    Y[7] = 0.25
    gid = np.arange(S) // 30
    sl = mo.score_series_batch(ref, Y)
    for g in (None, gid):
        for thr, topn, ml in ((0.0, 50, 10), (0.5, 20, 60), (0.0, 5000, 256)):
            a = co.batch_run(ref, Y, g, ml, topn, thr)
            b = mo.batch_run_arrays(ref, Y, g, ml, topn, thr, scores_lags=sl)
            assert len(a[0]) == len(b[0])
            np.testing.assert_allclose(a[0], b[0], rtol=0, atol=1e-12)
            assert (a[1] == b[1]).all() and (a[2] == b[2]).all()


def test_synthetic_generator_matches_the_library():
    """oracle/synth_gen.c (what bench.py's reference arm generates its rows with) against the library's host-side
    generator (muse_synth.cuh compiled for the host, itself checked against the device in test_gpu_parity.py)."""
    import muse_b200 as mb
    seed = 20261018
    for N in (480, 1440, 10080):
        rows = co.synth_rows(seed, 999_990, 24, N)
        for k in range(24):
            np.testing.assert_array_equal(rows[k], mb.synth_row(seed, 999_990 + k, N))
        np.testing.assert_array_equal(co.synth_reference(seed, N), mb.synth_reference(seed, N))
