// muse_xcorr.cuh -- the generic xCorr(x, y, n, normalize) of xcorr.go:102-153 for ANY length n
// (the reference's own tests use n = 5, xcorr_test.go:86-202), where the FFT kernels of
// muse_exact.cuh only exist for powers of two.
//
// The reference computes cc = irfft(rfft(xp) * conj(rfft(yp))) / n with LEADING zero pads
// (xcorr.go:70-80, :129-142); written out that is the circular correlation
//     cc[k] = sum_t xp[(t + k) mod n] * yp[t],      k = 0 .. n-1,
// times 1/(n-1) when both inputs were z-normalised (:139-140: 1/(n(n-1)) on gonum's unnormalised inverse,
// which is n times the correlation) and as is otherwise (:142: 1/n on the same).  This file evaluates the sum
// directly in fp64: n^2 fused multiply-adds spread over n/8 warps, inputs staged in shared memory
// tile by tile.  It is a utility, not a hot path: one pair per call.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace muse {

// One block per input row (blockIdx.x = 0: x, 1: y).  z-normalisation as xcorr.go:84-95: subtract the
// mean, sample (n-1) std of the centred row with gonum's corrected two-pass variance, multiply by 1/std;
// std == 0 (or a single sample) -> *std_zero = 1 and the caller returns (nil, 0, 0) as :109-126 does.
// The row lands at the END of the n-long padded output (leading zeros).
__global__ void __launch_bounds__(256)
xcorr_prepare_kernel(const double *__restrict__ x, long long x_len, const double *__restrict__ y, long long y_len,
                     long long n, int normalize, double *__restrict__ xp, double *__restrict__ yp, int *__restrict__ std_zero) {
    const double *src = blockIdx.x == 0 ? x : y;
    const long long len = blockIdx.x == 0 ? x_len : y_len;
    double *dst = blockIdx.x == 0 ? xp : yp;
    __shared__ double red[256];
    __shared__ double s_mean, s_scale;
    const int tid = threadIdx.x;
    auto block_sum = [&](double v) {
        red[tid] = v;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (tid < off) red[tid] += red[tid + off];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    double mean = 0.0, scale = 1.0;
    if (normalize) {
        double s = 0.0;
        for (long long i = tid; i < len; i += 256) s += src[i];
        mean = block_sum(s) / (double)len;
        double ss = 0.0, sd = 0.0;
        for (long long i = tid; i < len; i += 256) {
            const double d = src[i] - mean;
            ss += d * d;
            sd += d;
        }
        ss = block_sum(ss);
        sd = block_sum(sd);
        // the row handed to stat.StdDev is already centred: its own mean is sd/len (round-off), and the
        // corrected two-pass variance about THAT mean is (ss - sd^2/len)/(len-1) up to O(eps^2)
        const double var = (ss - sd * sd / (double)len) / (double)(len - 1);
        const double sdev = sqrt(var);
        if (tid == 0) {
            s_mean = mean;
            s_scale = 1.0 / sdev;
            if (!(sdev != 0.0) || len < 2) atomicExch(std_zero, 1);   // == 0 (NaN stays NaN like the reference: 1/NaN)
        }
        __syncthreads();
        mean = s_mean;
        scale = s_scale;
    }
    const long long pad = n - len;
    for (long long i = tid; i < n; i += 256) dst[i] = i < pad ? 0.0 : (normalize ? (src[i - pad] - mean) * scale : src[i - pad]);
}

// cc[k] = scale * sum_t xp[(t + k) mod n] * yp[t].  A block owns XC_LAGS consecutive lags; one warp per 8 lags
// would leave most of a 148-SM part idle at n = 5 and is irrelevant at that size, so the layout is chosen for
// long rows: thread = one lag, the t loop runs over shared-memory tiles of yp and the matching window of xp
// (tile + XC_LAGS - 1 samples, every element read from global memory once per block).
constexpr int XC_LAGS = 128;
constexpr int XC_TILE = 1024;
__global__ void __launch_bounds__(XC_LAGS)
xcorr_direct_kernel(const double *__restrict__ xp, const double *__restrict__ yp, long long n, double scale,
                    double *__restrict__ cc) {
    __shared__ double ys[XC_TILE];
    __shared__ double xs[XC_TILE + XC_LAGS];
    const long long k0 = (long long)blockIdx.x * XC_LAGS;
    const long long k = k0 + threadIdx.x;
    double acc = 0.0;
    for (long long t0 = 0; t0 < n; t0 += XC_TILE) {
        const int tl = (int)min((long long)XC_TILE, n - t0);
        for (int i = threadIdx.x; i < tl; i += XC_LAGS) ys[i] = yp[t0 + i];
        for (int i = threadIdx.x; i < tl + XC_LAGS - 1; i += XC_LAGS) xs[i] = xp[(t0 + k0 + i) % n];
        __syncthreads();
        if (k < n) {
            for (int i = 0; i < tl; i++) acc = fma(xs[i + threadIdx.x], ys[i], acc);
        }
        __syncthreads();
    }
    if (k < n) cc[k] = acc * scale;
}

// maxAbsIndex (xcorr.go:39-50): first index whose |cc| is STRICTLY larger than everything before it, starting
// from the value 0 at index 0 (all-zero rows give 0, NaNs never compare larger); then the wrap of :149-151.
// One block; lowest index wins ties.
__global__ void __launch_bounds__(256)
xcorr_argmax_kernel(const double *__restrict__ cc, long long n, long long *__restrict__ out_lag, double *__restrict__ out_val) {
    __shared__ double bv[256];
    __shared__ long long bi[256];
    double best = 0.0;
    long long idx = 0;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const double a = fabs(cc[i]);
        if (a > best) {      // NaN compares false, like the Go loop
            best = a;
            idx = i;
        }
    }
    bv[threadIdx.x] = best;
    bi[threadIdx.x] = idx;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double ov = bv[threadIdx.x + off];
            const long long oi = bi[threadIdx.x + off];
            if (ov > bv[threadIdx.x] || (ov == bv[threadIdx.x] && oi < bi[threadIdx.x])) {
                bv[threadIdx.x] = ov;
                bi[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long mi = bv[0] > 0.0 ? bi[0] : 0;
        *out_val = cc[mi];
        *out_lag = mi > n / 2 ? mi - n : mi;
    }
}

}  // namespace muse
