"""ctypes loader for the C oracle (oracle/muse_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmuse_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.muse_oracle_next_pow_of2.restype = C.c_int64
        L.muse_oracle_next_pow_of2.argtypes = [C.c_double]
        L.muse_oracle_xcorr_with_x.argtypes = [dp, dp, C.c_int64, dp, ip, dp]
        L.muse_oracle_score_all.argtypes = [dp, C.c_int64, dp, C.c_int64, C.c_int, dp, ip, C.c_int]
        L.muse_oracle_batch_run.argtypes = [dp, C.c_int64, dp, C.c_int64, ip, ip, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_double, C.c_int, dp, ip, ip, ip, C.c_int]
        L.muse_oracle_max_threads.restype = C.c_int
        L.muse_oracle_synth_rows.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, dp]
        L.muse_oracle_synth_rows.restype = None
        L.muse_oracle_synth_reference.argtypes = [C.c_uint64, C.c_int64, dp]
        L.muse_oracle_synth_reference.restype = None
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def synth_rows(seed: int, first: int, count: int, N: int) -> np.ndarray:
    """Rows first .. first+count-1 of the benchmark's synthetic store (synth_gen.c)."""
    out = np.empty((count, N))
    lib().muse_oracle_synth_rows(seed, first, count, N, _d(out))
    return out


def synth_reference(seed: int, N: int) -> np.ndarray:
    out = np.empty(N)
    lib().muse_oracle_synth_reference(seed, N, _d(out))
    return out


def max_threads() -> int:
    return int(lib().muse_oracle_max_threads())


def xcorr_with_x(ref, y):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    N = ref.size
    n = int(lib().muse_oracle_next_pow_of2(float(N)))
    cc = np.zeros(n)
    lag = C.c_int64(0)
    mv = C.c_double(0)
    rc = lib().muse_oracle_xcorr_with_x(_d(ref), _d(y), N, _d(cc), C.byref(lag), C.byref(mv))
    return rc, cc, int(lag.value), float(mv.value)


def score_all(ref, Y, signed=False, nthreads=0):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    S, N = Y.shape
    scores = np.zeros(S)
    lags = np.zeros(S, dtype=np.int64)
    rc = lib().muse_oracle_score_all(_d(ref), N, _d(Y), S, int(signed), _d(scores), _i(lags), nthreads)
    if rc:
        raise ValueError("Invalid input query, Standard deviation of zero")
    return scores, lags


def batch_run(ref, Y, group_ids, max_lag, top_n, threshold, sign_filter=0, nthreads=0):
    """group_ids: dense ids per series or None (each series its own group)."""
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    S, N = Y.shape
    if group_ids is None:
        members = np.arange(S, dtype=np.int64)
        goff = np.arange(S + 1, dtype=np.int64)
        G = S
    else:
        gid = np.asarray(group_ids, dtype=np.int64)
        G = int(gid.max()) + 1 if S else 0
        members = np.argsort(gid, kind="stable").astype(np.int64)
        goff = np.zeros(G + 1, dtype=np.int64)
        np.cumsum(np.bincount(gid, minlength=G), out=goff[1:])
    cap = max(1, min(int(top_n), G))
    sc = np.zeros(cap)
    lg = np.zeros(cap, dtype=np.int64)
    ix = np.zeros(cap, dtype=np.int64)
    n_out = C.c_int64(0)
    rc = lib().muse_oracle_batch_run(_d(ref), N, _d(Y), S, _i(members), _i(goff), G, max_lag, top_n,
                                     threshold, sign_filter, _d(sc), _i(lg), _i(ix), C.byref(n_out), nthreads)
    if rc:
        raise ValueError("Invalid input query, Standard deviation of zero")
    k = int(n_out.value)
    return sc[:k], lg[:k], ix[:k]
