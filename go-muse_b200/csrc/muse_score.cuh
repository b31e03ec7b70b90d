// muse_score.cuh -- per-thread phases of the fused score kernel (host+device).
//
// One series is handled by T = M/P cooperating threads (M = n/2 complex points of the
// half-length real-FFT trick, P points per thread).  The phases, in order, with a
// barrier between them in the kernel (and a plain loop over t in the CPU emulator):
//
//   load_row        128-bit coalesced loads of the fp64 row with the LEADING zero pad
//                   folded in (xcorr.go:176-181); partial sum for the mean
//   center_row      y -= mean (xcorr.go:85-86); partial sums for the sample std (:88)
//   fft passes      forward FFT_M (muse_fft.cuh)                    (xcorr.go:183)
//   pointwise_phase real split + conj(Y)*X + re-pack                (xcorr.go:184-185)
//   fft passes      inverse as swap(FFT(swap))                      (xcorr.go:186-187)
//   argmax_local    first index of max |cc| over ALL n lags         (xcorr.go:39-50,189)
//   finish          1/std scale, abs+clamp (muse_batch.go:74-77) or signed clamp
//                   (muse.go:72-76), wrap lag (xcorr.go:192-194), std==0 -> (0, 0)
//                   (xcorr.go:165-168)
#pragma once

#include "muse_fft.cuh"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include <math.h>

namespace muse {

typedef cx<double> cd;

// Streaming 16-byte load (read once, do not pollute L1).
MUSE_HD cd load_pair_stream(const double *p) {
#if defined(__CUDA_ARCH__)
    cd r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
#else
    return cd{p[0], p[1]};
#endif
}
MUSE_HD double load_one_stream(const double *p) {
#if defined(__CUDA_ARCH__)
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
#else
    return p[0];
#endif
}

// z[j] = (ypad[2j], ypad[2j+1]) for j = t + r*T; ypad = leading zeros ++ row.
// EVEN_N: N (hence pad) is even and the row is 16-byte aligned -> vector loads.
template <int LOG2M, int LOG2P, bool EVEN_N>
MUSE_HD void load_row(cd *v, const double *row, int N, int t, double &sum) {
    using G = Geo<LOG2M, LOG2P>;
    const int n = 2 * G::M;
    const int pad = n - N;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int r = 0; r < G::P; r++) {
        const int i0 = 2 * (t + r * G::T) - pad;
        cd val{0.0, 0.0};
        if (EVEN_N) {
            if (i0 >= 0) val = load_pair_stream(row + i0);
        } else {
            if (i0 >= 0) val.x = load_one_stream(row + i0);
            if (i0 + 1 >= 0) val.y = load_one_stream(row + i0 + 1);
        }
        v[r] = val;
        s0 += val.x;
        s1 += val.y;
    }
    sum = s0 + s1;
}

template <int LOG2M, int LOG2P>
MUSE_HD void center_row(cd *v, int N, int t, double mu, double &ss, double &comp) {
    using G = Geo<LOG2M, LOG2P>;
    const int n = 2 * G::M;
    const int pad = n - N;
    double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int r = 0; r < G::P; r++) {
        const int i0 = 2 * (t + r * G::T) - pad;
        if (i0 >= 0) {
            const double d = v[r].x - mu;
            v[r].x = d;
            a0 += d * d;
            c0 += d;
        }
        if (i0 + 1 >= 0) {
            const double d = v[r].y - mu;
            v[r].y = d;
            a1 += d * d;
            c1 += d;
        }
    }
    ss = a0 + a1;
    comp = c0 + c1;
}

// Pairs (k, M-k), k = t, t+T, ... <= M/2.  twn[k] = exp(-2*pi*i*k/n); Xt = X/(2n), M+1 entries.
template <int LOG2M, int LOG2P, typename F, typename TW>
MUSE_HD void pointwise_phase(cx<F> *sm, int t, const TW *Xt, const TW *twn) {
    using G = Geo<LOG2M, LOG2P>;
    for (int k = t; k <= G::M / 2; k += G::T) {
        const int m = (G::M - k) & (G::M - 1);
        const cx<F> zk = sm[G::pad(k)], zm = sm[G::pad(m)];
        const TW w = twn[k], xk = Xt[k], xm = Xt[G::M - k];
        cx<F> ok, om;
        pointwise_pair(zk, zm, cx<F>{(F)w.x, (F)w.y}, xk, xm, ok, om);
        sm[G::pad(m)] = om;
        sm[G::pad(k)] = ok;
    }
}

struct Peak {
    double a;   // |value|; 0 when nothing strictly positive was seen
    double v;   // signed value
    int idx;    // cc index (0..n-1)
};

MUSE_HD void peak_merge(Peak &p, double a, double v, int idx) {
    // strictly greater wins; equal magnitude -> lowest index (maxAbsIndex scans upward, xcorr.go:42-47)
    if (a > p.a || (a == p.a && a > 0.0 && idx < p.idx)) {
        p.a = a;
        p.v = v;
        p.idx = idx;
    }
}

// After the last inverse pass v holds swap(IFFT): cc[2e] = v.y, cc[2e+1] = v.x.
template <int LOG2M, int LOG2P, typename F>
MUSE_HD Peak argmax_local(const cx<F> *v, int t) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr int PASS = G::NPASS - 1;
    constexpr int LR = G::log2r(PASS);
    constexpr int R = 1 << LR;
    constexpr int NB = G::P / R;
    Peak pk{0.0, 0.0, 0x7fffffff};
#pragma unroll
    for (int c = 0; c < NB; c++) {
#pragma unroll
        for (int j = 0; j < R; j++) {
            const cx<F> val = v[c * R + Perm<R>::at(j)];
            const int e = last_pass_index<LOG2M, LOG2P>(t, c, j);
            peak_merge(pk, fabs((double)val.y), (double)val.y, 2 * e);
            peak_merge(pk, fabs((double)val.x), (double)val.x, 2 * e + 1);
        }
    }
    return pk;
}

// Final per-series result from the reduced peak and the reduced sums.
MUSE_HD void finish_series(Peak pk, double ss, double comp, int N, int n, bool signed_scores,
                           double &score, int &lag) {
    const double var = (ss - comp * comp / (double)N) / (double)(N - 1);
    const double sd = sqrt(var);
    if (sd == 0.0) {   // xcorr.go:165-168
        score = 0.0;
        lag = 0;
        return;
    }
    int mi = pk.a > 0.0 ? pk.idx : 0;
    double mv = (pk.a > 0.0 ? pk.v : 0.0) * (1.0 / sd);
    if (signed_scores) {
        if (mv > 1.0) mv = 1.0;
        else if (mv < -1.0) mv = -1.0;
    } else {
        mv = fabs(mv);
        if (mv > 1.0) mv = 1.0;
    }
    if (mi > n / 2) mi -= n;
    score = mv;
    lag = mi;
}

}  // namespace muse
