// muse_screen.cuh -- fp32 screening pass: a rigorous UPPER BOUND on every series' score.
//
// For the score of xcorr.go:160-197 / muse_batch.go:74-77,
//     score = min(1, max_k |cc[k]|),   cc[k] = (1/n) * sum_f C_f * exp(+2*pi*i*f*k/n),
//     C_f = conj(Y_f) * X_f,
// the triangle inequality gives  max_k |cc[k]| <= (1/n) * sum_f |Y_f| * |X_f|  =: U.
// The warp and sub-warp kernels (n <= 2048) bound U itself once more, by Cauchy-Schwarz on each mirror pair (f, n/2 - f)
// of the half-length transform: |2Y_k| A[k] + |2Y_(M-k)| A[M-k] <= sqrt(|2Y_k|^2 + |2Y_(M-k)|^2) sqrt(A[k]^2 + A[M-k]^2), and
// the first factor is sqrt(2 (|e|^2 + |d|^2)) straight from e = Z_k + conj(Z_(M-k)), d = Z_k - conj(Z_(M-k)) -- no split
// twiddle, one square root per pair.  Still an upper bound on the score, a little looser (it passes 5.5 % of the
// benchmark's series to the second stage instead of 3.6 %), a third cheaper in the loop every series runs.
// U needs only the FORWARD transform of the series, no inverse, no arg-max.  The kernel
// computes U in fp32 (forward FFT_M of the half-length packing, split, |Y_f|, dot with the
// precomputed |X_f| weights) and adds a slack that covers every fp32 rounding in the
// chain, so that U >= exact fp64 score holds for every series.  Batch.Run then sends only
// series whose U can still reach the top-N cut-off through the exact fp64 kernel
// (muse_exact.cuh); its results are therefore identical to scoring everything exactly.
//
// Error budget (all in units of the normalised score, |cc| <= 1):
//   * input: y - pivot is formed in fp64, then rounded to fp32: |err| <= 2^-24 * 2*max|y-mean|
//     per sample, and max|y-mean| <= sqrt(N-1)*std, so the induced cc error is
//     <= ||x'||_2 * sqrt(N) * 2^-23 * sqrt(N-1) * ... <= 2^-23 * sqrt(N) ~ 4.5e-6 at N = 1440;
//   * fp32 FFT (radix-32 x radix-32): relative l2 error of the spectrum <~ 10 * 2^-24 ~ 6e-7,
//     which moves U by at most ||X||_2*||dY||_2/n <= 6e-7;
//   * mean, std, magnitude and accumulation roundings: each <= ~1e-6 relative.
// The sum stays below 2e-5 for N <= 16384; SCREEN_SLACK = 1e-4 leaves a 5x margin over that budget and 450x over the
// worst error observed on the adversarial generator at every FFT length (2.2e-7, profiles/screen_error_survey.json).
// The slack is what sends series to the exact kernel: every series whose score lies within 2 slacks of the top-N cut-off
// is re-scored in fp64 (at 2e-4, rounds 1 and 2a: 4.0 k of 1 M series at C3, 777 k of 256 M pairs = 11.7 of 62 ms at C5).
// tests/test_gpu_screen.py checks U >= exact on adversarial inputs (offsets of 1e9, spikes,
// 1e-12 and 1e+12 amplitudes, trends) and records the smallest observed margin.
#pragma once

#include <vector_functions.h>   // float4 / make_float4 (host builds of the emulators too)
#include <vector_types.h>

#include "muse_score.cuh"

namespace muse {

typedef cx<float> cf;

#ifndef MUSE_SCREEN_SLACK
#define MUSE_SCREEN_SLACK 1e-4f
#endif
// warp kernel: sample variance window.  std >= 1e-10: a flushed magnitude (< 1.1e-19) is below
// 1.1e-9 in units of the score per bin; std <= 1e14: |2Y_k|^2 <= (4*N*std)^2 * N < 3e38.
#define MUSE_SCREEN_VAR_MIN 1e-20f
#define MUSE_SCREEN_VAR_MAX 1e28f
// |mean| <= 1e8 * std: the fp64 sum of N values of size |mean| is uncertain by at most
// N * 2^-53 * N * |mean| (sequential worst case), i.e. the mean by 1.6e-13 * |mean| at N = 1440; a
// shift of the mean by d moves a score by at most d / std, so the limit keeps the disagreement
// between this kernel's mean and the exact kernel's (summed in another order) below 2e-5, a tenth of
// the slack.  Rows beyond it are flagged once per store by row_offset_flags_kernel (a register for
// the mean across the FFT costs this kernel 5 %) and always go to the exact kernel.
#define MUSE_SCREEN_OFFSET_MAX 1e8
// running cut-off: lower bounds are counted in MUSE_CUT_BINS = 64^3 bins of [0, 1] (3.8e-6 wide, far below the
// slack), with 64^2 and 64 group counters above them (MUSE_CUT_WORDS counters in all)
#define MUSE_CUT_BINS 262144
#define MUSE_CUT_WORDS (64 + 64 * 64 + MUSE_CUT_BINS)

// Per-row statistics of the store, computed once per appended row (the batched mean/std kernel of the
// z-normalisation, xcorr.go:84-95): the fp64 mean, and 1/std in fp32 -- or NaN when no fp32 statement may be
// made about the row (constant row, variance outside [MUSE_SCREEN_VAR_MIN, MUSE_SCREEN_VAR_MAX], non-finite
// samples, |mean| > MUSE_SCREEN_OFFSET_MAX * std): the NaN propagates into the bound and sends the series to
// the exact kernel.
struct alignas(16) RowStat {
    double mean;
    float rstd;
    unsigned pad;
};

struct ScreenParams {
    const double *slab;
    int64_t ld;
    int64_t count;
    int N;
    const cf *twp;        // per-pass twiddles (fill_pass_twiddles), fp32
    const float4 *sw;     // (w_k.x, w_k.y, A[k], A[M-k]) for k < M/2: split twiddle exp(-2*pi*i*k/n), weights |X|/(2n)*(1|2) rounded up
    const float *sb;      // warp and sub-warp kernels: B[k] = sqrt(2 (A[k]^2 + A[M-k]^2)) rounded up, k < M/2 (the pair bound below)
    float a_mid;          // A[M/2]
    // ---- fused second stage ----
    const float4 *sx;     // (Xt[k].x, Xt[k].y, Xt[M-k].x, Xt[M-k].y) in fp32, k < M/2 (Xt = X/(2n))
    cf x_mid;             // Xt[M/2]
    float *out_L;         // [count] certain lower bound on a score that certainly passes the lag filter, else -1
    const RowStat *row_stat;          // [count] fp64 mean and fp32 1/std of each row (row_stats_kernel, once per appended row)
    unsigned *cut_bits;   // running lower bound on the final top-N cut-off (float bits, only ever raised)
    unsigned *cut_hist;   // [MUSE_CUT_WORDS] 64 coarse, 64*64 middle, 64^3 fine counts of lower bounds
    unsigned long long *n_refined;
    int top_n;            // series needed above the cut before it may rise
    int win_lo, win_len;  // lag window in the kernel's rotated cc index: (idx - win_lo) mod n <= win_len
    float thr;            // threshold rounded up to fp32 (a lower bound must reach it to count)
    int grouped;          // grouped run: every series is refined, bounds ignore the window, no cut-off
    float *out_U;         // [count] upper bound on the score (already clamped to <= 1 + slack)
    // ---- muse_screen_big.cuh ----
    const cf *twi;        // fp32 twiddles of the transposed inverse (fill_big_inverse_twiddles)
    const int64_t *slot_of;           // grouped runs: group-table slot of every series
    unsigned long long *group_L;      // grouped runs: [slots] running best lower bound of each group (float bits, low word)
    signed char *out_W;   // grouped runs: [count] 1 = peak certainly inside the lag window, -1 = certainly outside, 0 = undecided
};

// ---- maxima of |cc'| inside / outside the lag window (rotated index: (idx - win_lo) mod n <= win_len) ----
// A register pair r holds (cc'[idx + 1], cc'[idx]).  The pairs of one ROW of the transform's output (one register slot
// over the threads of a series) cover row_len consecutive lags starting at `off`; whether any of them can be inside
// the window depends on the row alone, i.e. on kernel parameters: the test runs on the uniform datapath and all but the
// two or three rows the window touches take one 3-input max per pair instead of two compares, two masks and four
// selects.  Used by the block kernel of muse_screen_big.cuh (C4: 43.0 -> 42.1 ms ungrouped, 46.1 -> 45.0 ms grouped); in the
// warp kernels the 32 uniform branches cost more than they save (C3 2.10 -> 2.17 ms, C5's second stages 39.3 -> 42.3 ms).
MUSE_HD bool window_row_hit(int off, int row_len, int win_lo, int win_len, int n) {
    const int d = (off - win_lo) & (n - 1);
    return d <= win_len || d > n - row_len;
}
MUSE_HD void window_pair(cx<float> r, int idx, int win_len, int nmask, bool hit, float &m_in, float &m_out) {
    const float a0 = fabsf(r.y), a1 = fabsf(r.x);
    if (hit) {
        const bool in0 = (idx & nmask) <= win_len;
        const bool in1 = ((idx + 1) & nmask) <= win_len;
        m_in = fmaxf(m_in, fmaxf(in0 ? a0 : 0.f, in1 ? a1 : 0.f));
        m_out = fmaxf(m_out, fmaxf(in0 ? 0.f : a0, in1 ? 0.f : a1));
    } else {
        m_out = fmaxf(m_out, fmaxf(a0, a1));
    }
}

#if defined(__CUDACC__)

template <int T>
__device__ __forceinline__ float group_sum_f(float x) {
#pragma unroll
    for (int off = T / 2; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

// ---- cp.async.bulk / mbarrier plumbing of the warp kernel (SASS UBLKCP, SYNCS) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// sqrt.approx.ftz.f32 (one MUFU): 2 ulp, covered by the 1e-5 relative slack of the bound.  A
// subnormal |2Y_k|^2 is flushed to 0; the variance window MUSE_SCREEN_VAR_MIN/MAX keeps every
// bin that matters at the 1e-9 level far away from both ends of the fp32 range.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ void mbar_test(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Warp-collective: count the certain lower bound L of a series that certainly passes the filter and
// try to raise the running cut-off to the top_n-th largest lower bound counted so far.  All counts
// only grow, so every value read is a lower bound on the true count and any cut-off derived from
// them stays valid; lanes walk the counters from the top down, three levels of 64.
// One level of the walk: lanes read the 64 counters of a group from the top down (two per lane), find the
// counter in which the running count reaches `need`, and return its index within the group (-1: the
// group does not hold enough); `need` is reduced by what lies above that counter.
__device__ __forceinline__ int cut_level(const unsigned *cnt, int t, unsigned &need) {
    const unsigned h = ld_relaxed_u32(&cnt[63 - 2 * t]), l = ld_relaxed_u32(&cnt[62 - 2 * t]);
    unsigned incl = h + l;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
        if (t >= off) incl += o;
    }
    const unsigned excl = incl - (h + l);
    const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
    if (!hit) return -1;
    const int leader = __ffs(hit) - 1;
    const bool upper = excl + h >= need;
    const int idx = upper ? 63 - 2 * t : 62 - 2 * t;
    const unsigned above = upper ? excl : excl + h;
    need -= __shfl_sync(0xffffffffu, above, leader);
    return __shfl_sync(0xffffffffu, idx, leader);
}

__device__ __forceinline__ void cut_count_and_raise(const ScreenParams &prm, float L, int t) {
    int bin = (int)(L * (float)MUSE_CUT_BINS);
    bin = bin < 0 ? 0 : (bin >= MUSE_CUT_BINS ? MUSE_CUT_BINS - 1 : bin);
    unsigned *coarse = prm.cut_hist, *mid = coarse + 64, *fine = mid + 64 * 64;
    if (t == 0) {
        atomicAdd(&fine[bin], 1u);
        atomicAdd(&mid[bin >> 6], 1u);
        atomicAdd(&coarse[bin >> 12], 1u);
    }
    __syncwarp();
    unsigned need = (unsigned)prm.top_n;
    const int c = cut_level(coarse, t, need);
    if (c < 0) return;
    const int m = cut_level(mid + c * 64, t, need);
    if (m < 0) return;                  // the three levels are incremented separately: a reader may see them out of step
    const int f = cut_level(fine + (c * 64 + m) * 64, t, need);
    if (f < 0) return;
    if (t == 0) atomicMax(prm.cut_bits, __float_as_uint((float)((c * 64 + m) * 64 + f) / (float)MUSE_CUT_BINS));
}

// Refined decision for one series from the fp32 maxima of |cc| inside / outside the lag window
// (both already divided by std): returns the new upper bound (-1: certainly outside the window,
// results.go:46-48 drops the series) and sets L when the peak is certainly inside.
// |fp32 cc - exact cc| <= MUSE_SCREEN_SLACK / 2 in score units (error budget above, with the inverse
// transform doubling the FFT term); every decision leaves a full slack.
// grouped != 0: the window is not this series' business (its group's representative is decided by
// the scores alone, muse_batch.go:87-89): upper and lower bound on the score itself.
__device__ __forceinline__ float refine_decide(float U, float s_in, float s_out, float &L, int grouped) {
    if (!(s_in == s_in) || !(s_out == s_out)) return U;
    const float u32 = fminf(fmaxf(s_in, s_out) * 1.00001f, 1.f) + MUSE_SCREEN_SLACK;
    if (grouped) {
        L = fmaxf(fminf(fmaxf(s_in, s_out) * 0.99999f, 1.f) - MUSE_SCREEN_SLACK, 0.f);
        return fminf(U, u32);
    }
    if (s_out * 0.99999f - MUSE_SCREEN_SLACK > s_in * 1.00001f + MUSE_SCREEN_SLACK) return -1.f;
    if (s_in * 0.99999f - MUSE_SCREEN_SLACK > s_out * 1.00001f + MUSE_SCREEN_SLACK)
        L = fminf(s_in * 0.99999f, 1.f) - MUSE_SCREEN_SLACK;      // certainly inside
    return fminf(U, u32);
}

#if defined(__CUDACC__)
// RowStat from the sums s1 = sum (y - pivot), s2 = sum (y - pivot)^2 of a row of N samples.
__device__ __forceinline__ RowStat make_row_stat(double pivot, double s1, double s2, int N) {
    const double mean = pivot + s1 / N;
    const double var = (s2 - s1 * s1 / N) / (N - 1);
    const bool ok = var >= (double)MUSE_SCREEN_VAR_MIN && var <= (double)MUSE_SCREEN_VAR_MAX &&
                    mean * mean <= MUSE_SCREEN_OFFSET_MAX * MUSE_SCREEN_OFFSET_MAX * var;      // false for NaN / Inf
    RowStat r;
    r.mean = mean;
    r.rstd = ok ? (float)(1.0 / sqrt(var)) : __int_as_float(0x7fc00000);
    r.pad = 0u;
    return r;
}
#endif

// RowStat of rows first .. first+count-1; for an odd series length the row's mean is also written into the first pad
// column of the row (ld >= N + 1 then), so that the screening kernels, which read rows as pairs of samples, see a
// centred value of exactly 0 there whatever the ingest left in it.  One warp per row; sums are taken about the row's first sample so
// that the one-pass variance does not cancel (|row[0] - mean| <= sqrt(N-1) * std, so the cancellation costs at
// most a factor N of the 1e-16).
static __global__ void row_stats_kernel(double *__restrict__ slab, int64_t ld, int N, int64_t first, int64_t count,
                                 RowStat *__restrict__ stat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < count; i += nwarp) {
        const double *row = slab + (first + i) * ld;
        const double pivot = row[0];
        double s1 = 0.0, s2 = 0.0;
        for (int k = lane; k < N; k += 32) {
            const double x = row[k] - pivot;
            s1 += x;
            s2 = fma(x, x, s2);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if (lane == 0) {
            const RowStat r = make_row_stat(pivot, s1, s2, N);
            stat[first + i] = r;
            if (N & 1) slab[(first + i) * ld + N] = r.mean;
        }
    }
}

struct ScreenWarpCfg {
    using G = Geo<10, 5>;                       // M = 1024 = 32 x 32: one warp per series
    static constexpr int MAX_WARPS = 12;        // 3 warps per scheduler at up to 168 registers per thread (16 at 128 measured equal)
    static constexpr int NREF = 2;              // exchange buffers for the (rare) second stage, shared under a lock
    static constexpr size_t SMEM_BUDGET = 227 * 1024;
    static constexpr size_t EX_BYTES = ((size_t)(G::MP + 1) * sizeof(cf) + 127) / 128 * 128;   // padded FFT exchange buffer
    static size_t row_bytes(int N) { const size_t r = ((size_t)N * 8 + 127) / 128 * 128; return r > EX_BYTES ? r : EX_BYTES; }
    static size_t warp_bytes(int N) { return row_bytes(N); }
    static int warps(int N) {
        const size_t w = (SMEM_BUDGET - NREF * EX_BYTES) / warp_bytes(N);
        return (int)(w > MAX_WARPS ? MAX_WARPS : w);
    }
    static size_t smem_bytes(int N) { return (size_t)warps(N) * warp_bytes(N) + NREF * EX_BYTES; }
    static int nz(int N) { return ((N + 1) / 2 + 31) / 32; }
};

__device__ __forceinline__ int ex_acquire(unsigned *locks, int t, int hint) {
    int slot = hint & (ScreenWarpCfg::NREF - 1);
    while (true) {
        unsigned old = 1u;
        if (t == 0) old = atomicCAS(&locks[slot], 0u, 1u);
        if (__shfl_sync(0xffffffffu, old, 0) == 0u) break;
        slot = (slot + 1) & (ScreenWarpCfg::NREF - 1);
    }
    return slot;
}
__device__ __forceinline__ void ex_release(unsigned *locks, int t, int slot) {
    __syncwarp();
    if (t == 0) {
        __threadfence_block();
        atomicExch(&locks[slot], 0u);
    }
}

// Persistent kernel, one block per SM, every warp an independent pipeline over its own series
// (pos = first, first + stride, ...): its row buffer is refilled by the next cp.async.bulk as
// soon as the 23 x 16 bytes per lane are in registers, so the DRAM latency of row i+1 hides
// behind the FFT of row i and ~10 rows per SM are in flight at any time.  The FFT exchange
// goes through a separate per-warp buffer.
//
// NZ = ceil(N/64) (17..32 for n = 2048) is a template parameter so that the loads, the
// conversions and the first radix-4 stage of the zero rows disappear at compile time.  The
// |Y_f| bound is invariant under rotation of the padded series, so the zeros trail.
//
// Split: lane t holds Z[t + 32j] in slot j.  Bins k and M-k share e = Z[k] + conj(Z[M-k]) and
// w*o, 2Y[k] = e + w*o, 2conj(Y[M-k]) = e - w*o, so each lane walks only j < 16 (k < 512) and
// gets both magnitudes from one mirror exchange: Z[M-k] sits in lane (32-t)%32, slot 31-j
// (lane 0: its own slot (32-j)%32; k = 0 pairs with itself and yields the DC and Nyquist
// terms).  Only k = 512 (lane 0, slot 16, its own mirror) is left over: |2Y[512]| = 2|Z[512]|.
//
// Fused second stage: U is loose (it ignores every phase), and on the benchmark's data
// a few percent of the series have U above the top-N cut-off.  A warp whose U reaches the
// RUNNING cut-off therefore goes on, with the spectrum still in registers: conj(Y)*X in fp32,
// inverse FFT, max |cc| inside and outside the lag window.  That replaces U by the much
// tighter u32 = |cc|_max/std + slack (or by -1 when the peak is certainly outside the window:
// the series fails results.go:46-48), and yields a certain LOWER bound for series that certainly
// pass the filter.  Lower bounds are counted in a global histogram; the top_n-th largest so far
// is a valid lower bound on the final cut-off and becomes the new running cut-off (it only ever
// rises, so a stale read is merely less sharp).  Afterwards only series with out_U >= final cut
// -- a few hundred -- need the exact fp64 kernel.  The cc index is rotated by pad = n - N against
// xcorr.go's (zeros trail here, lead there): true index = (idx - pad) mod n.
template <int NZ>
__global__ void __launch_bounds__(ScreenWarpCfg::MAX_WARPS * 32, 1)
score_screen_warp_kernel(const ScreenParams prm, const unsigned warp_bytes, const unsigned row_bytes) {
    using C = ScreenWarpCfg;
    using G = typename C::G;
    constexpr int P = 32, M = G::M;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[C::MAX_WARPS];
    __shared__ unsigned ex_locks[C::NREF];

    const int w = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    // 32-bit series indices (a store holds < 2^31 series): the kernel sits on a register cliff at 168
    const int count = (int)prm.count;
    const int stride = (int)(gridDim.x * (blockDim.x >> 5));
    const int pos0 = (int)(blockIdx.x * (blockDim.x >> 5)) + w;
    unsigned char *buf = smem_raw + (size_t)w * warp_bytes;
    const cd *rowc = reinterpret_cast<const cd *>(buf);
    cf *sm = reinterpret_cast<cf *>(buf);                       // forward exchange: the row buffer itself
    unsigned char *refbase = smem_raw + (size_t)(blockDim.x >> 5) * warp_bytes;
    const int N = prm.N;
    const int Nh = (N + 1) >> 1;      // complex slots holding samples (odd N: the pad column of the last one holds the row's mean, RowStat)
    const unsigned bar = smem_u32(&bars[w]);
    const int partner = (P - t) & (P - 1);
    const bool lane0 = (t == 0);
    const bool last_in = t + (NZ - 1) * 32 < Nh;      // is this lane's slot of the last row a sample?

    if (threadIdx.x < C::NREF) ex_locks[threadIdx.x] = 0u;
    if (t == 0) {
        mbar_init(bar, 1);
        if (pos0 < count) bulk_load(smem_u32(buf), prm.slab + (int64_t)pos0 * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
    }
    __syncthreads();

    // mbarrier phase parity = iteration parity; it rides in bit 31 of the loop counter (count < 2^31):
    // a separate loop-carried register is one too many for the allocator at 168 and gets spilled
    for (unsigned pp = (unsigned)pos0; (int)(pp & 0x7fffffffu) < count; pp = ((pp & 0x7fffffffu) + (unsigned)stride) | (~pp & 0x80000000u)) {
        const int pos = (int)(pp & 0x7fffffffu);
        const unsigned phase = pp >> 31;
        // running cut-off: one lane reads it (so that the whole warp takes the same branch below) at
        // the top of the iteration; the value is consumed only after U is known, a thousand
        // instructions later, so the L2 round trip stays off the critical path.  +inf = no refinement
        unsigned cut_raw = 0u;
        if (t == 0) cut_raw = ld_relaxed_u32(prm.cut_bits);
        const RowStat rs = prm.row_stat[pos];   // fp64 mean and 1/std from the ingest pass (xcorr.go:84-95); in flight with the row
        const double mu = rs.mean;
        mbar_wait(bar, phase);      // every lane waits itself (one polling lane + __syncwarp measured 3x slower)

        // ---- centred samples in fp32 straight from the row buffer ----
        cf v[P];
#pragma unroll
        for (int r = 0; r < P; r++) {
            if (r < NZ) {
                const cd x = (r == NZ - 1 && !last_in) ? cd{mu, mu} : rowc[t + r * 32];
                v[r] = cf{(float)(x.x - mu), (float)(x.y - mu)};
            } else {
                v[r] = cf{0.f, 0.f};
            }
        }
        __syncwarp();                       // the row is consumed: the buffer becomes the exchange buffer
        const int next = pos + stride;      // < 2^31: the launcher keeps count + stride below it

        // ---- forward FFT_1024: pruned radix-32, twiddle, exchange through smem, radix-32 ----
        Dft32Lead<NZ, float>::run(v);
#pragma unroll
        for (int j = 0; j < P; j++) {
            cf val = v[Perm<P>::at(j)];
            if (j > 0) val = cmul(val, prm.twp[(j - 1) * 32 + t]);          // W_1024^(j*t)
            sm[G::pad(32 * t + j)] = val;
        }
        __syncwarp();
        fft_pass_load<10, 5, 1, float>(v, sm, t);
        __syncwarp();
        if (t == 0 && next < count) {       // the exchange is over: hand the buffer to the copy engine for the next row
            asm volatile("" ::"r"(__float_as_uint(v[P - 1].y)) : "memory");
            // the generic-proxy reads and writes of the exchange are ordered before the async-proxy write of the copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(smem_u32(buf), prm.slab + (int64_t)next * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
        }
        Dft<P, float>::run(v);                          // v[Perm(j)] = Z[t + 32*j]

        // ---- the bound over the mirror pairs (k, M-k), k = t + 32*j, j < 16 ----
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < P / 2; j++) {
            const cf zk = v[Perm<P>::at(j)];
            const cf zp = v[Perm<P>::at(P - 1 - j)];
            const cf zs = v[Perm<P>::at((P - j) & (P - 1))];
            cf src, zm;
            src.x = lane0 ? zs.x : zp.x;
            src.y = lane0 ? zs.y : zp.y;
            zm.x = __shfl_sync(0xffffffffu, src.x, partner);
            zm.y = __shfl_sync(0xffffffffu, src.y, partner);
            // pair bound: 2Y_k = e + w o and 2conj(Y_(M-k)) = e - w o with |w| = 1 give |2Y_k|^2 + |2Y_(M-k)|^2 = 2 (|e|^2 + |d|^2),
            // d = Z_k - conj(Z_(M-k)) (|d| = |o|), so by Cauchy-Schwarz
            //   |2Y_k| A[k] + |2Y_(M-k)| A[M-k] <= sqrt(|e|^2 + |d|^2) * B[k]:
            // no twiddle, one square root per pair.  It passes 5.5 % of the benchmark's series on to the second stage where
            // the bin-by-bin sum passed 3.6 %, for a third fewer instructions in this loop on every series.
            const cf zmc = cconj(zm);
            const cf e = cadd(zk, zmc);
            const cf d = csub(zk, zmc);
            const cf q = pfma(d, d, pmul(e, e));
            acc = fmaf(sqrt_approx(q.x + q.y), prm.sb[t + 32 * j], acc);
        }
        {   // k = 512: lane 0, slot 16 (weight 0 on the other lanes)
            const cf z = v[Perm<P>::at(P / 2)];
            const cf q = pmul(z, z);
            acc = fmaf(sqrt_approx(q.x + q.y), lane0 ? 2.f * prm.a_mid : 0.f, acc);
        }
        acc = group_sum_f<32>(acc);

        // rstd is NaN for a row no fp32 statement may be made about (RowStat); NaN/Inf samples make acc NaN:
        // either way the bound is NaN and the exact kernel decides
        float U = acc * rs.rstd * 1.00001f + MUSE_SCREEN_SLACK;
        if (!(U == U)) U = 2.f;
        float L = -1.f;
        // The broadcast must not be scheduled before this point: it would wait for the load
        // issued at the top of the iteration (measured: +8 % kernel time).  Its source lane
        // therefore depends on acc, which exists only now; it is 0 unless acc has one particular
        // NaN pattern, and then lane 1's zeros merely send the series through the refinement.
        const int bsrc = (__float_as_uint(acc) == 0x7fc12345u) ? 1 : 0;
        const float cut_now = __uint_as_float(__shfl_sync(0xffffffffu, cut_raw, bsrc));
        if (U >= cut_now && U < 1.5f) {      // warp-uniform: U comes out of a butterfly reduction
            // ---- conj(Y)*X on the mirror pairs of the split (pointwise_pair), in place:
            //      Z'[k] -> own slot j;  Z'[M-k] -> the partner's slot 31-j (lane 0: its own slot 32-j) ----
#pragma unroll
            for (int j = 0; j < P / 2; j++) {
                const cf zk = v[Perm<P>::at(j)];
                const cf zp = v[Perm<P>::at(P - 1 - j)];
                const cf zs = v[Perm<P>::at((P - j) & (P - 1))];
                cf src, zm;
                src.x = lane0 ? zs.x : zp.x;
                src.y = lane0 ? zs.y : zp.y;
                zm.x = __shfl_sync(0xffffffffu, src.x, partner);
                zm.y = __shfl_sync(0xffffffffu, src.y, partner);
                const float4 s = prm.sw[t + 32 * j];
                const float4 x = prm.sx[t + 32 * j];
                cf ok, om;
                pointwise_pair(zk, zm, cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                cf rcv;
                rcv.x = __shfl_sync(0xffffffffu, om.x, partner);
                rcv.y = __shfl_sync(0xffffffffu, om.y, partner);
                v[Perm<P>::at(j)] = ok;
                cf &hi = v[Perm<P>::at(P - 1 - j)];
                hi.x = lane0 ? hi.x : rcv.x;
                hi.y = lane0 ? hi.y : rcv.y;
                if (j > 0) {
                    cf &own = v[Perm<P>::at(P - j)];
                    own.x = lane0 ? om.x : own.x;
                    own.y = lane0 ? om.y : own.y;
                }
            }
            {   // k = 512 pairs with itself (lane 0, slot 16); w_512 = -i
                cf &mid = v[Perm<P>::at(P / 2)];
                cf ok, om;
                pointwise_pair(mid, mid, cf{0.f, -1.f}, prm.x_mid, prm.x_mid, ok, om);
                mid.x = lane0 ? ok.x : mid.x;
                mid.y = lane0 ? ok.y : mid.y;
            }
            // ---- inverse FFT_1024 as swap(FFT(swap(.))) (pointwise_pair stores the swapped values) ----
            cf u[P];
#pragma unroll
            for (int j = 0; j < P; j++) u[j] = v[Perm<P>::at(j)];
            Dft<P, float>::run(u);
            {
                const int slot = ex_acquire(ex_locks, t, w);
                cf *smr = reinterpret_cast<cf *>(refbase + (size_t)slot * C::EX_BYTES);
#pragma unroll
                for (int j = 0; j < P; j++) {
                    cf val = u[Perm<P>::at(j)];
                    if (j > 0) val = cmul(val, prm.twp[(j - 1) * 32 + t]);
                    smr[G::pad(32 * t + j)] = val;
                }
                __syncwarp();
                fft_pass_load<10, 5, 1, float>(u, smr, t);
                ex_release(ex_locks, t, slot);
            }
            Dft<P, float>::run(u);      // u[Perm(j)] = (cc'[2i+1], cc'[2i]), i = t + 32*j; cc' = std * cc rotated by pad
            // ---- max |cc'| inside and outside the lag window ----
            float m_in = 0.f, m_out = 0.f;
            const int base = 2 * t - prm.win_lo;
#pragma unroll
            for (int j = 0; j < P; j++) {
                const cf r = u[Perm<P>::at(j)];
                const bool in0 = ((base + 64 * j) & (2 * M - 1)) <= prm.win_len;
                const bool in1 = ((base + 64 * j + 1) & (2 * M - 1)) <= prm.win_len;
                const float a0 = fabsf(r.y), a1 = fabsf(r.x);
                m_in = fmaxf(m_in, in0 ? a0 : 0.f);
                m_out = fmaxf(m_out, in0 ? 0.f : a0);
                m_in = fmaxf(m_in, in1 ? a1 : 0.f);
                m_out = fmaxf(m_out, in1 ? 0.f : a1);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, off));
                m_out = fmaxf(m_out, __shfl_xor_sync(0xffffffffu, m_out, off));
            }
            const float rstd = rs.rstd;
            const float s_in = m_in * rstd, s_out = m_out * rstd;
            U = refine_decide(U, s_in, s_out, L, prm.grouped);
            if (t == 0) atomicAdd(prm.n_refined, 1ull);
            // ---- a certain pass at or above the running cut-off: count it and try to raise the cut-off ----
            if (L >= prm.thr && L >= cut_now) cut_count_and_raise(prm, L, t);
        }
        if (t == 0) {
            prm.out_U[pos] = U;
            if (prm.out_L) prm.out_L[pos] = L;
        }
    }
}

#endif  // __CUDACC__

}  // namespace muse
