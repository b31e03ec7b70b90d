"""CPU oracle for go-muse's Batch.Run hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the reference algorithm.  It is the
checker for the CUDA path, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.

Every function cites the reference file:line (relative to /root/reference) it
follows.  Arithmetic that the reference delegates to un-vendored Go modules
(gonum.org/v1/gonum v0.7.0: dsp/fourier, floats, stat; go.mod:5-9) is restated
from the published algorithms:
  * floats.Sum       -> sequential left-to-right fp64 sum
  * stat.StdDev      -> sqrt of the corrected two-pass unbiased (n-1) variance
                        (Chan/Golub/LeVeque eq. 1.7: (ss - comp^2/n)/(n-1))
  * fourier.FFT      -> FFTPACK real FFT; Coefficients returns n/2+1 complex
                        values, Sequence is the UN-normalised inverse.  Here
                        numpy's pocketfft stands in; any correct fp64 FFT agrees
                        to ~1e-15, far inside the 1e-9 parity gate.
Parity pinning: the reference cannot be executed in this image (no Go
toolchain), so the oracle is pinned against every expectation in the
reference's own tests (tests/golden/reference_kats.json, transcribed from
xcorr_test.go, muse_batch_test.go, muse_test.go, group_test.go, labels_test.go);
see tests/test_oracle.py.  Beyond those KATs (N>480, top-N eviction, threshold>0,
multi-label grouping) parity is unpinned and the oracle is the only authority.

Map-iteration order: Go iterates maps in random order (group.go:83,
muse_batch.go:24); the oracle iterates in insertion order, which is one of the
orders the reference can take.  Ties therefore resolve to the FIRST inserted.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

SIGN_FILTER_POS = 1   # results.go:23
SIGN_FILTER_NEG = -1  # results.go:24
SIGN_FILTER_ANY = 0   # results.go:25
DEFAULT_LABEL = "uid"  # labels.go:7


# --------------------------------------------------------------------------
# xcorr.go
# --------------------------------------------------------------------------
def next_pow_of2(val: float) -> int:
    """xcorr.go:19-24 -- int(pow(2, ceil(log(val)/log(2)))); <=0 -> 0."""
    if val <= 0:
        return 0
    return int(math.pow(2.0, math.ceil(math.log(val) / math.log(2))))


def max_abs_index(x: np.ndarray) -> int:
    """xcorr.go:39-50 -- first index whose |v| is STRICTLY larger than every
    earlier one (start value 0, so all-zero / all-NaN input gives 0)."""
    a = np.abs(x)
    if a.size == 0:
        return 0
    # NaN never compares greater (math.Abs(v) > math.Abs(maxVal) is false)
    a = np.where(np.isnan(a), -1.0, a)
    mi = int(np.argmax(a))  # argmax returns the first maximal index
    if not a[mi] > 0.0:
        return 0
    return mi


def zero_pad(x: np.ndarray, n: int) -> np.ndarray:
    """xcorr.go:70-80 -- LEADING zeros; n < len(x) returns x itself."""
    x = np.asarray(x, dtype=np.float64)
    if n < x.size:
        return x
    out = np.zeros(n, dtype=np.float64)
    out[n - x.size:] = x
    return out


def _seq_sum(x: np.ndarray) -> float:
    """gonum floats.Sum: sequential accumulation (cumsum is sequential)."""
    if x.size == 0:
        return 0.0
    return float(np.cumsum(x, dtype=np.float64)[-1])


def std_dev(x: np.ndarray) -> float:
    """gonum stat.StdDev(x, nil): sqrt of corrected two-pass unbiased variance."""
    n = x.size
    mean = _seq_sum(x) / n
    d = x - mean
    ss = _seq_sum(d * d)
    comp = _seq_sum(d)
    with np.errstate(invalid="ignore", divide="ignore"):
        var = (ss - comp * comp / n) / np.float64(n - 1)
    return float(np.sqrt(var))


def z_normalize(x: np.ndarray) -> Optional[np.ndarray]:
    """xcorr.go:84-95 -- subtract mean, divide by SAMPLE std; std==0 -> None
    (errStdDevZero).  Returns a new array (the reference works in place, F4)."""
    x = np.array(x, dtype=np.float64, copy=True)
    n = float(x.size)
    x += -_seq_sum(x) / n            # floats.AddConst(-floats.Sum(x)/n, x)
    s = std_dev(x)                   # stat.StdDev(x, nil)
    if s == 0:
        return None
    x *= 1.0 / s                     # floats.Scale(1/stdX, x)
    return x


def _coefficients(seq: np.ndarray) -> np.ndarray:
    """fourier.FFT.Coefficients: n real -> n/2+1 complex."""
    return np.fft.rfft(seq)


def _sequence(coef: np.ndarray, n: int) -> np.ndarray:
    """fourier.FFT.Sequence: UN-normalised inverse real FFT."""
    return np.fft.irfft(coef, n) * n


def _wrap(mi: int, n: int) -> int:
    """xcorr.go:149-151 / 192-194 -- mi > n/2 (integer division) -> mi - n."""
    if mi > n // 2:
        mi -= n
    return mi


def x_corr(x, y, n: int, normalize: bool):
    """xcorr.go:102-153 -- generic cross correlation. Returns (cc, lag, value);
    (None, 0, 0.0) when a std is zero."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = max(n, x.size, y.size)                      # :104-106
    if normalize:
        x = z_normalize(x)
        if x is None:
            return None, 0, 0.0
        y = z_normalize(y)
        if y is None:
            return None, 0, 0.0
    xp = zero_pad(x, n)
    yp = zero_pad(y, n)
    X = _coefficients(xp)
    Y = _coefficients(yp)
    cc = _sequence(X * np.conj(Y), n)               # conj(Y); mult(X, Y)
    if normalize:
        cc = cc * (1.0 / float(n * (n - 1)))        # :139-140
    else:
        cc = cc * (1.0 / float(n))                  # :142
    mi = max_abs_index(cc)
    mv = float(cc[mi])
    return cc, _wrap(mi, n), mv


def ref_spectrum(ref) -> Tuple[Optional[np.ndarray], int]:
    """muse_batch.go:35-47 (== muse.go:27-39) -- n = nextPowOf2(N);
    x' = znorm(ref)/(N-1), leading-zero pad to n, X = rfft(x').
    Returns (None, n) when std(ref)==0 ("Invalid input query")."""
    ref = np.asarray(ref, dtype=np.float64)
    n = next_pow_of2(float(ref.size))
    x = z_normalize(ref)
    if x is None:
        return None, n
    x = x * (1.0 / float(x.size - 1))
    return _coefficients(zero_pad(x, n)), n


def x_corr_with_x(X: np.ndarray, y, n: int):
    """xcorr.go:160-197 -- cross correlation against a precomputed X.
    Returns (cc, lag, value); (None, 0, 0.0) when std(y)==0 (:165-168)."""
    yn = z_normalize(np.asarray(y, dtype=np.float64))
    if yn is None:
        return None, 0, 0.0
    seq = np.zeros(n, dtype=np.float64)             # :176-181 leading zeros
    seq[n - yn.size:] = yn
    C = np.conj(_coefficients(seq)) * X             # :183-185
    cc = _sequence(C, n) * (1.0 / float(n))         # :186-187
    mi = max_abs_index(cc)                          # :189
    mv = float(cc[mi])
    return cc, _wrap(mi, n), mv


def tie_lags(cc: Optional[np.ndarray], n: int, tol: float = 1e-9) -> List[int]:
    """All lags whose |cc| is within tol of the series' max |cc| (SURVEY F4:
    the reference's own tests pin a round-off-level tie, so a parity check has to
    accept any member of this set)."""
    if cc is None:
        return [0]
    a = np.abs(cc)
    m = np.nanmax(a) if a.size else 0.0
    return [_wrap(int(i), n) for i in np.nonzero(a >= m - tol)[0]]


# --------------------------------------------------------------------------
# labels.go / series.go / group.go
# --------------------------------------------------------------------------
class Labels:
    """labels.go:14-73."""

    def __init__(self, labels: Dict[str, str]):
        self.labels = dict(labels)
        self.keys = sorted(self.labels.keys())      # :25-29 (byte-wise sort)

    def __len__(self):
        return len(self.labels)

    def Len(self):
        return len(self.labels)

    def Keys(self):
        return self.keys

    def Get(self, key):
        if key in self.labels:
            return self.labels[key], True
        return "", False

    def ID(self, labels: Optional[List[str]] = None) -> str:
        """labels.go:54-73 -- "k1:v1,k2:v2" over the SORTED requested keys,
        absent keys skipped; sorts the caller's list in place."""
        if not labels:
            labels = self.keys
        else:
            labels.sort()
        out = ""
        for k in labels:
            if k in self.labels:
                out += k + ":" + self.labels[k] + ","
        return out[:-1] if out.endswith(",") else out


_uid_counter = [0]


class Series:
    """series.go:8-42.  Unlike the reference the values are never mutated."""

    def __init__(self, y, labels: Optional[Labels] = None):
        if labels is None or labels.Len() == 0:
            _uid_counter[0] += 1                    # stands in for uuid4 (:17)
            labels = Labels({DEFAULT_LABEL: "oracle-uid-%d" % _uid_counter[0]})
        self.y = np.asarray(y, dtype=np.float64)
        self.labels = labels

    def Length(self):
        return int(self.y.size)

    def Values(self):
        return self.y

    def Labels(self):
        return self.labels

    def UID(self):
        return self.labels.ID(list(self.labels.Keys()))


class Group:
    """group.go:7-104."""

    def __init__(self, name: str):
        self.Name = name
        self.n = 0
        self.index: Dict[str, List[str]] = {}
        self.registry: Dict[str, Series] = {}       # insertion-ordered

    def Length(self):
        return self.n

    def Add(self, *series: Series):
        """group.go:31-56 -- raises ValueError where the reference returns an error."""
        for s in series:
            if len(s.labels.Keys()) == 0:
                raise ValueError("Invalid Series with no labels")
            uid = s.UID()
            if uid in self.registry:
                raise ValueError("Series with label:values, %s, already exists within group, %s"
                                 % (uid, self.Name))
            if len(self.registry) == 0:
                self.n = s.Length()
            elif s.Length() != self.n:
                raise ValueError("Timeseries has length %d, but current group has length %d"
                                 % (s.Length(), self.n))
            self.registry[uid] = s

    def FilterByLabelValues(self, labels: Labels) -> List[Series]:
        """group.go:60-71."""
        guid = labels.ID(list(labels.Keys()))
        if guid in self.index:
            return [self.registry[u] for u in self.index[guid]]
        return []

    def indexLabelValues(self, group_by: Optional[List[str]]) -> List[Labels]:
        """group.go:76-104 (including the :86-88 quirk: with no labels given,
        groupByLabels is overwritten with the first visited series' keys)."""
        distinct: List[Labels] = []
        self.index = {}
        group_by = list(group_by) if group_by else []
        for uid, s in self.registry.items():
            if len(group_by) != 0:
                guid = s.labels.ID(group_by)
            else:
                guid = uid
                group_by = list(s.labels.Keys())
            if guid not in self.index:
                lv = {}
                for name in group_by:
                    v, ok = s.labels.Get(name)
                    if ok:
                        lv[name] = v
                distinct.append(Labels(lv))
            self.index.setdefault(guid, []).append(uid)
        return distinct


# --------------------------------------------------------------------------
# scores.go / results.go
# --------------------------------------------------------------------------
@dataclass
class Score:
    """scores.go:11-15."""
    Labels: Optional[Labels] = None
    Lag: int = 0
    PercentScore: float = 0.0


class _Heap:
    """container/heap over Scores with Less = |a| < |b| (scores.go:25-27);
    up/down follow Go's container/heap so tie behaviour is the reference's."""

    def __init__(self):
        self.s: List[Score] = []

    def _less(self, i, j):
        return abs(self.s[i].PercentScore) < abs(self.s[j].PercentScore)

    def _up(self, j):
        while True:
            i = (j - 1) // 2
            if i == j or j <= 0 or not self._less(j, i):
                break
            self.s[i], self.s[j] = self.s[j], self.s[i]
            j = i

    def _down(self, i0, n):
        i = i0
        while True:
            j1 = 2 * i + 1
            if j1 >= n or j1 < 0:
                break
            j = j1
            j2 = j1 + 1
            if j2 < n and self._less(j2, j1):
                j = j2
            if not self._less(j, i):
                break
            self.s[i], self.s[j] = self.s[j], self.s[i]
            i = j
        return i > i0

    def push(self, x: Score):
        self.s.append(x)
        self._up(len(self.s) - 1)

    def pop(self) -> Score:
        n = len(self.s) - 1
        self.s[0], self.s[n] = self.s[n], self.s[0]
        self._down(0, n)
        return self.s.pop()

    def __len__(self):
        return len(self.s)


class Results:
    """results.go:11-87."""

    def __init__(self, max_lag: int, top_n: int, threshold: float, sign_filter: int):
        self.MaxLag = max_lag
        self.TopN = top_n
        self.Threshold = threshold
        self.SignFilter = sign_filter
        self.scores = _Heap()

    def passed(self, s: Score) -> bool:
        """results.go:46-52."""
        return (abs(float(s.Lag)) <= float(self.MaxLag)
                and abs(s.PercentScore) >= self.Threshold
                and (self.SignFilter == SIGN_FILTER_ANY
                     or (s.PercentScore > 0 and self.SignFilter == SIGN_FILTER_POS)
                     or (s.PercentScore < 0 and self.SignFilter == SIGN_FILTER_NEG)))

    def Update(self, s: Score):
        """results.go:55-72 -- replace the heap root only if STRICTLY greater."""
        if s.Labels is None:
            return
        if self.passed(s):
            if len(self.scores) == self.TopN:
                if self.TopN > 0 and abs(s.PercentScore) > abs(self.scores.s[0].PercentScore):
                    self.scores.pop()
                    self.scores.push(s)
            else:
                self.scores.push(s)

    def Fetch(self) -> Tuple[List[Score], float]:
        """results.go:75-87 -- drains the heap; DESCENDING |score|; mean |score|
        (NaN when empty: 0/0)."""
        num = len(self.scores)
        out: List[Optional[Score]] = [None] * num
        total = 0.0
        for i in range(num - 1, -1, -1):
            sc = self.scores.pop()
            total += abs(sc.PercentScore)
            out[i] = sc
        mean = total / num if num else float("nan")
        return out, mean  # type: ignore[return-value]


# --------------------------------------------------------------------------
# muse_batch.go / muse.go
# --------------------------------------------------------------------------
class Batch:
    """muse_batch.go:13-130."""

    def __init__(self, ref: Series, comp: Group, results: Results, cc: int = 1):
        for uid, s in comp.registry.items():                       # :24-28
            if ref.Length() != s.Length():
                raise ValueError("%s from comparison group series does not have the same "
                                 "length as the reference" % uid)
        if cc < 1:
            cc = 1
        X, n = ref_spectrum(ref.Values())                          # :35-47
        if X is None:
            raise ValueError("Invalid input query, Standard deviation of zero")
        self.n = n
        self.x = X
        self.Comparison = comp
        self.Results = results
        self.Concurrency = cc

    def _score_single(self, label_values: Labels) -> Score:
        """muse_batch.go:56-93."""
        max_score = Score()
        for ts in self.Comparison.FilterByLabelValues(label_values):
            _, lag, mv = x_corr_with_x(self.x, ts.Values(), self.n)
            mv = abs(mv)                                           # :74
            if mv > 1.0:                                           # :75-77
                mv = 1.0
            comp = Score(Labels=ts.Labels(), Lag=lag, PercentScore=mv)
            if comp.PercentScore > max_score.PercentScore or max_score.Labels is None:  # :87
                max_score = comp
        return max_score

    def Run(self, group_by: Optional[List[str]] = None):
        """muse_batch.go:99-130."""
        for lv in self.Comparison.indexLabelValues(group_by):
            self.Results.Update(self._score_single(lv))
        return None


class Muse:
    """muse.go:15-92 -- the signed single-group path."""

    def __init__(self, ref: Series, results: Results):
        if ref.Length() < 1:
            raise ValueError("Reference series length must be greater than zero")
        X, n = ref_spectrum(ref.Values())
        if X is None:
            raise ValueError("Invalid input query, Standard deviation of zero")
        self.refN = ref.Length()
        self.n = n
        self.x = X
        self.Results = results

    def Run(self, comp: Sequence[Series]):
        if len(comp) == 0:                                         # :47-50
            return None
        max_score = Score()
        for ts in comp:
            if ts.Length() != self.refN:                           # :68-70
                raise ValueError("Encountered a comparison graph with differing length "
                                 "than the reference")
            _, lag, mv = x_corr_with_x(self.x, ts.Values(), self.n)
            if mv > 1.0:                                           # :72-76
                mv = 1.0
            elif mv < -1.0:
                mv = -1.0
            comp_s = Score(Labels=ts.Labels(), Lag=lag, PercentScore=mv)
            if abs(comp_s.PercentScore) > abs(max_score.PercentScore) or max_score.Labels is None:
                max_score = comp_s
        self.Results.Update(max_score)                             # :90
        return None


# --------------------------------------------------------------------------
# Vectorised array form of the same path (for parity at sizes where the
# object-per-series form above would take minutes).  tests/test_oracle.py
# checks it against the scalar form.
# --------------------------------------------------------------------------
def score_series_batch(ref, Y: np.ndarray, signed: bool = False, chunk: int = 4096,
                       want_ties: bool = False, tie_tol: float = 1e-9):
    """Per-series (score, lag) for every row of Y[S, N] against `ref`.

    Same arithmetic as x_corr_with_x + muse_batch.go:74-77 (abs, clamp to 1) or,
    with signed=True, muse.go:72-76 (clamp to [-1, 1]); std==0 rows give (0, 0)
    (xcorr.go:165-168).  Returns (scores f64[S], lags i64[S]) and, when
    want_ties, a list of per-series tie-lag lists.
    """
    Y = np.asarray(Y, dtype=np.float64)
    S, N = Y.shape
    X, n = ref_spectrum(ref)
    if X is None:
        raise ValueError("Invalid input query, Standard deviation of zero")
    scores = np.zeros(S, dtype=np.float64)
    lags = np.zeros(S, dtype=np.int64)
    ties: List[List[int]] = []
    for s0 in range(0, S, chunk):
        y = Y[s0:s0 + chunk]
        m = y.shape[0]
        mean = np.cumsum(y, axis=1)[:, -1] / float(N)
        d = y - mean[:, None]
        # stat.StdDev on the centred data: its own mean, then corrected two-pass
        mean2 = np.cumsum(d, axis=1)[:, -1] / float(N)
        e = d - mean2[:, None]
        ss = np.cumsum(e * e, axis=1)[:, -1]
        comp = np.cumsum(e, axis=1)[:, -1]
        with np.errstate(invalid="ignore", divide="ignore"):
            std = np.sqrt((ss - comp * comp / float(N)) / np.float64(N - 1))
            z = d * (1.0 / std)[:, None]
        seq = np.zeros((m, n), dtype=np.float64)
        seq[:, n - N:] = z
        C = np.conj(np.fft.rfft(seq, axis=1)) * X[None, :]
        cc = np.fft.irfft(C, n, axis=1) * n * (1.0 / float(n))
        a = np.abs(cc)
        a_cmp = np.where(np.isnan(a), -1.0, a)
        mi = np.argmax(a_cmp, axis=1)
        amax = a_cmp[np.arange(m), mi]
        mi = np.where(amax > 0.0, mi, 0)
        mv = cc[np.arange(m), mi]
        zero = std == 0
        lag = np.where(mi > n // 2, mi - n, mi)
        if signed:
            sc = np.clip(mv, -1.0, 1.0)
        else:
            sc = np.abs(mv)
            sc = np.where(sc > 1.0, 1.0, sc)
        sc = np.where(zero, 0.0, sc)
        lag = np.where(zero, 0, lag)
        scores[s0:s0 + m] = sc
        lags[s0:s0 + m] = lag
        if want_ties:
            for r in range(m):
                if zero[r]:
                    ties.append([0])
                else:
                    ties.append(tie_lags(cc[r], n, tie_tol))
    if want_ties:
        return scores, lags, ties
    return scores, lags


def batch_run_arrays(ref, Y: np.ndarray, group_ids: Optional[np.ndarray], max_lag: int,
                     top_n: int, threshold: float, sign_filter: int = SIGN_FILTER_ANY,
                     scores_lags=None):
    """Array form of Batch.Run + Results.Fetch for ONE Run on a fresh Results.

    group_ids[S] are dense group numbers in first-appearance order (None: each
    series its own group, muse_batch.go:96-98).  Follows muse_batch.go:87-89
    (group max BEFORE the filter, first member wins ties), results.go:46-52
    (filter on the representative) and :55-87 (top-N by |score|, descending).
    Ties at the top-N boundary resolve to the group visited first, as
    results.go:62-66 (strictly-greater replace) does for that visiting order.
    Returns (scores, lags, series_idx) of the kept representatives, descending.
    """
    if scores_lags is None:
        scores, lags = score_series_batch(ref, Y)
    else:
        scores, lags = scores_lags
    S = scores.shape[0]
    if group_ids is None:
        rep = np.arange(S, dtype=np.int64)
    else:
        gid = np.asarray(group_ids, dtype=np.int64)
        G = int(gid.max()) + 1 if S else 0
        best = np.full(G, -1, dtype=np.int64)
        for i in range(S):                                         # first member, then strict >
            g = gid[i]
            b = best[g]
            if b < 0 or scores[i] > scores[b]:
                best[g] = i
        rep = best[best >= 0]
    sc = scores[rep]
    lg = lags[rep]
    with np.errstate(invalid="ignore"):
        ok = (np.abs(lg) <= max_lag) & (np.abs(sc) >= threshold)
        if sign_filter == SIGN_FILTER_POS:
            ok &= sc > 0
        elif sign_filter == SIGN_FILTER_NEG:
            ok &= sc < 0
    rep, sc, lg = rep[ok], sc[ok], lg[ok]
    order = np.argsort(-np.abs(sc), kind="stable")                 # stable: first visited first
    order = order[:max(top_n, 0)]
    return sc[order], lg[order], rep[order]
