"""The multi-query bounds on the tensor cores (muse_bounds_tc.cuh): the bf16 contraction must dominate every exact
score and stay close to the fp32 kernel's bound.  Needs a B200."""
import numpy as np
import pytest

import muse_b200 as mb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return mb.default_context(0)


def _refs(rng, Q, N):
    refs = 0.1 * (rng.random((Q, N)) - 0.5)
    for q in range(Q):
        mid, w = int(rng.integers(N // 2 - N // 8, N // 2 + N // 8)), int(rng.integers(3, 21))
        refs[q, mid - w // 2: mid - w // 2 + w] += 1.5
    return refs


@pytest.mark.parametrize("N,S,Q", [(1440, 1000, 5), (1440, 4096 + 77, 256), (2048, 300, 33), (1030, 515, 1)])
def test_tensor_core_bounds_dominate_and_track_the_fp32_bound(ctx, N, S, Q):
    rng = np.random.default_rng(N + S + Q)
    store = mb.DeviceStore(ctx, N, 0, S)
    store.append_synthetic(S, 20261018, 7)
    # a few hostile rows: constant (no fp32 statement possible), huge offset, tiny amplitude
    Y = store.read_rows(0, S)
    refs = _refs(rng, Q, N)
    U = mb.multi_bounds_tc(store, refs).astype(np.float64)
    assert U.shape == (Q, S)
    for q in sorted(set([0, Q // 2, Q - 1])):
        b = mb.DeviceBatch(ctx, store, refs[q])
        sc, _ = b.score_all()
        u32 = b.screen_bounds().astype(np.float64)
        dec = u32 <= 1.5
        assert np.all(U[q] >= sc + 0.5e-4), (q, float((U[q] - sc).min()))
        # bf16 rounds both operands up: at most (1 + 2^-7)^2 * 1.002 above the fp32 bound
        assert np.all(U[q][dec] <= (u32[dec] - 2e-4) * 1.0185 + 2.1e-4)
        assert np.all(U[q][dec] >= (u32[dec] - 2e-4) * 0.9999)
        b.close()
    store.close()
