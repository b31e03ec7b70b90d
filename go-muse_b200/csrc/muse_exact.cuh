// muse_exact.cuh -- the exact fp64 fused score kernel (device side).
//
// One launch = for every series of the slab (or of an index list): z-normalise,
// zero-pad, real FFT, conj(Y)*X, inverse FFT, max-|cc|/argmax over all n lags, abs +
// clamp -> (score, lag).  It is xCorrWithX + the per-member part of scoreSingle
// (xcorr.go:160-197, muse_batch.go:73-77) with nothing but the row read from and
// 12 bytes written to HBM per series.
#pragma once

#include "muse_score.cuh"

namespace muse {

enum { MODE_SCORE = 0, MODE_REF = 1, MODE_CC = 2 };

struct ExactParams {
    const double *slab;      // [rows][ld] fp64
    int64_t ld;              // row stride in doubles (multiple of 16)
    int64_t count;           // series to process (an upper bound when count_ptr is set)
    const unsigned long long *count_ptr;   // optional: actual count in device memory (no host round trip)
    const int32_t *idx;      // optional gather list (local row numbers), else NULL
    int N;                   // series length
    int signed_scores;
    const cd *Xt;            // X/(2n), M+1 entries
    const cd *twM;           // exp(-2*pi*i*k/M), M entries
    const cd *twn;           // exp(-2*pi*i*k/n), M/2+1 entries
    double *out_score;       // MODE_SCORE: [rows]   MODE_CC: cc[n]   MODE_REF: unused
    int32_t *out_lag;        // MODE_SCORE: [rows]
    cd *out_X;               // MODE_REF: Xt out (M+1)
    int32_t *out_flag;       // MODE_REF / MODE_CC: 1 when std == 0
    const ExactParams *batch;   // BATCH instantiations: the parameters of query blockIdx.y (device memory); everything above unused
};

template <int LOG2M, int LOG2P>
struct ExactCfg {
    using G = Geo<LOG2M, LOG2P>;
    static constexpr int T = G::T;
    static constexpr int TB = T > 128 ? T : 128;       // threads per block
    static constexpr int SPB = TB / T;                 // series per block
    static constexpr int NW = T > 32 ? T / 32 : 1;     // warps per series
    static constexpr int SM_ELEMS = G::MP + 1;         // padded elements per series
    static constexpr size_t SMEM = (size_t)SPB * SM_ELEMS * sizeof(cd);
};

#if defined(__CUDACC__)

template <int T>
__device__ __forceinline__ void group_sync() {
    if (T > 32) __syncthreads();
    else __syncwarp();
}

// Sum over the T threads of a series.  red: [warps per block] scratch.
template <int T>
__device__ __forceinline__ double group_sum(double x, double *red, int sib) {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (T > 32) {
        constexpr int NW = T / 32;
        const int wib = threadIdx.x >> 5;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[wib] = x;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red[sib * NW + w];
        x = s;
    }
    return x;
}

template <int T>
__device__ __forceinline__ Peak group_peak(Peak p, double *red, int sib) {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, p.a, off);
        const double v = __shfl_xor_sync(0xffffffffu, p.v, off);
        const int i = __shfl_xor_sync(0xffffffffu, p.idx, off);
        peak_merge(p, a, v, i);
    }
    if (T > 32) {
        constexpr int NW = T / 32;
        const int wib = threadIdx.x >> 5;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            red[wib * 3 + 0] = p.a;
            red[wib * 3 + 1] = p.v;
            red[wib * 3 + 2] = (double)p.idx;
        }
        __syncthreads();
        Peak r{0.0, 0.0, 0x7fffffff};
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const int o = (sib * NW + w) * 3;
            peak_merge(r, red[o], red[o + 1], (int)red[o + 2]);
        }
        p = r;
    }
    return p;
}

template <int LOG2M, int LOG2P, int PASS, bool FROM_REGS>
__device__ __forceinline__ void fft_to_last(cd *v, cd *sm, int t, const cd *__restrict__ twM) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr bool LAST = (PASS == G::NPASS - 1);
    if (!FROM_REGS) {
        fft_pass_load<LOG2M, LOG2P, PASS, double>(v, sm, t);
        group_sync<G::T>();
    }
    if constexpr (!LAST) {
        fft_pass_compute_store<LOG2M, LOG2P, PASS, double, cd>(v, sm, t, twM);
        group_sync<G::T>();
        fft_to_last<LOG2M, LOG2P, PASS + 1, false>(v, sm, t, twM);
    }
}

// BATCH: one launch serves many queries, blockIdx.y picks the query's parameter block (muse_multi_run's reference
// preparation and tails: one launch per stage instead of one per query)
template <int LOG2M, int LOG2P, int MODE, int MINB = 1, bool BATCH = false>
__global__ void __launch_bounds__(ExactCfg<LOG2M, LOG2P>::TB, MINB)
score_exact_kernel(const ExactParams prm0) {
    const ExactParams prm = BATCH ? prm0.batch[blockIdx.y] : prm0;
    using G = Geo<LOG2M, LOG2P>;
    using C = ExactCfg<LOG2M, LOG2P>;
    constexpr int T = C::T, M = G::M, n = 2 * M;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[48];

    const int sib = threadIdx.x / T;
    const int t = threadIdx.x - sib * T;
    const int64_t pos = (int64_t)blockIdx.x * C::SPB + sib;
    const int64_t count = prm.count_ptr ? (int64_t)*prm.count_ptr : prm.count;
    if ((int64_t)blockIdx.x * C::SPB >= count) return;      // whole block beyond the list (uniform)
    const bool valid = pos < count;
    const int64_t cpos = valid ? pos : count - 1;
    const int64_t row = prm.idx ? (int64_t)prm.idx[cpos] : cpos;
    const double *rowp = prm.slab + row * prm.ld;
    cd *sm = reinterpret_cast<cd *>(smem_raw) + (size_t)sib * C::SM_ELEMS;
    const int N = prm.N;

    cd v[G::P];
    double sum;
    if ((N & 1) == 0) load_row<LOG2M, LOG2P, true>(v, rowp, N, t, sum);
    else load_row<LOG2M, LOG2P, false>(v, rowp, N, t, sum);
    sum = group_sum<T>(sum, red, sib);
    const double mu = sum / (double)N;          // xcorr.go:85-86
    double ss, comp;
    center_row<LOG2M, LOG2P>(v, N, t, mu, ss, comp);

    // forward FFT_M; the last pass lands in smem in natural order
    fft_to_last<LOG2M, LOG2P, 0, true>(v, sm, t, prm.twM);
    fft_pass_compute_store<LOG2M, LOG2P, G::NPASS - 1, double, cd>(v, sm, t, prm.twM);
    group_sync<T>();

    ss = group_sum<T>(ss, red, sib);
    comp = group_sum<T>(comp, red, sib);
    const double var = (ss - comp * comp / (double)N) / (double)(N - 1);   // stat.StdDev, xcorr.go:88
    const double sd = sqrt(var);

    if (MODE == MODE_REF) {
        // muse_batch.go:38-47: X = rfft(zeroPad(znorm(ref)/(N-1), n)), stored as X/(2n)
        if (t == 0 && sib == 0) *prm.out_flag = (sd == 0.0) ? 1 : 0;
        const double scale = 1.0 / (4.0 * (double)n * sd * (double)(N - 1));
        if (sib == 0) {
            for (int k = t; k <= M / 2; k += T) {
                const int m = (M - k) & (M - 1);
                cd a, b;
                untangle_pair(sm[G::pad(k)], sm[G::pad(m)], prm.twn[k], scale, a, b);
                prm.out_X[M - k] = b;
                prm.out_X[k] = a;
            }
        }
        return;
    }

    pointwise_phase<LOG2M, LOG2P, double, cd>(sm, t, prm.Xt, prm.twn);
    group_sync<T>();

    // inverse as swap(FFT(swap)); the last pass stays in registers
    fft_to_last<LOG2M, LOG2P, 0, false>(v, sm, t, prm.twM);
    {
        constexpr int PASS = G::NPASS - 1;
        constexpr int R = 1 << G::log2r(PASS);
#pragma unroll
        for (int c = 0; c < G::P / R; c++) Dft<R, double>::run(v + c * R);
    }

    if (MODE == MODE_CC) {
        if (t == 0 && sib == 0) *prm.out_flag = (sd == 0.0) ? 1 : 0;
        if (sib == 0) {
            constexpr int PASS = G::NPASS - 1;
            constexpr int R = 1 << G::log2r(PASS);
            const double inv = sd == 0.0 ? 0.0 : 1.0 / sd;
#pragma unroll
            for (int c = 0; c < G::P / R; c++)
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int e = last_pass_index<LOG2M, LOG2P>(t, c, j);
                    const cd val = v[c * R + Perm<R>::at(j)];
                    prm.out_score[2 * e] = val.y * inv;
                    prm.out_score[2 * e + 1] = val.x * inv;
                }
        }
        return;
    }

    Peak pk = argmax_local<LOG2M, LOG2P, double>(v, t);
    pk = group_peak<T>(pk, red, sib);
    if (t == 0 && valid) {
        double score;
        int lag;
        finish_series(pk, ss, comp, N, n, prm.signed_scores != 0, score, lag);
        prm.out_score[row] = score;
        prm.out_lag[row] = lag;
    }
}

#endif  // __CUDACC__

}  // namespace muse
