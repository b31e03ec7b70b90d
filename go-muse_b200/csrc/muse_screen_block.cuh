// muse_screen_block.cuh -- the fp32 screening + fused refinement pass for FFT lengths above 2048
// (n = 4096, 8192, 16384: series of 2050 .. 16384 samples, e.g. one week at one sample per minute)
// and below it (n = 512, 1024: series of 258 .. 1024 samples).
//
// Same mathematics and the same contract as score_screen_warp_kernel (muse_screen.cuh): a rigorous
// upper bound U = (1/n) sum_f |Y_f||X_f| / std + slack for every series; a series whose U reaches the
// running top-N cut-off goes on to conj(Y)*X, the inverse transform and the maxima of |cc| inside and
// outside the lag window, which tighten the bound and feed the cut-off histogram.
//
// What differs is the shape: one series is M = n/2 complex points = 32 per thread on T = M/32
// threads (T = 256 at n = 16384), one series per block, so the Stockham passes (radix 32, 32, and
// 8 / 4 / 2) exchange through shared memory with block barriers, and the mirror pairs of the real
// split are read from shared memory instead of being shuffled.
//   * loads: each thread reads its 16-byte pairs straight from the slab (consecutive threads,
//     consecutive addresses); the NEXT row of the block is pulled into L2 by one
//     cp.async.bulk.prefetch.L2 at the top of the iteration, so the DRAM latency is paid under the
//     previous series' FFT and the loads themselves hit L2;
//   * centring: the fp64 mean and 1/std of every row are computed once at ingest (row_stats_kernel), so the
//     samples are converted as (float)(y - mean) on the way in and no block reduction precedes the transform.
#pragma once

#include "muse_screen.cuh"

namespace muse {

// LOG2M -> points per thread: 32 for M >= 2048, 16 for M = 512 (n = 1024), 8 for M = 256 (n = 512), so that a
// block is never smaller than a warp.  The transform is radix P x P x R_LAST (R_LAST = M / P^2).
template <int LOG2M>
struct ScreenBlockCfg {
    static_assert((LOG2M >= 11 && LOG2M <= 13) || LOG2M == 9 || LOG2M == 8, "block kernel: n = 512, 1024, 4096 .. 16384");
    static constexpr int LOG2P = LOG2M >= 11 ? 5 : (LOG2M == 9 ? 4 : 3);
    using G = Geo<LOG2M, LOG2P>;
    static constexpr int P = G::P;
    static constexpr int T = G::T;                   // threads per series == block size (32 .. 256)
    static constexpr int NWARP = T / 32;
    static constexpr size_t SMEM = (size_t)(G::MP + 1) * sizeof(cf);
    static constexpr int LAST = G::NPASS - 1;        // NPASS == 3
    static constexpr int LR_LAST = G::log2r(LAST);
    static constexpr int R_LAST = 1 << LR_LAST;
    static constexpr int NB_LAST = P / R_LAST;
    static constexpr int LS_LAST = 2 * LOG2P;
    static constexpr int TP = T + (T >> LOG2P);      // pad(t + T*x) = pad(t) + TP*x: T is a multiple of P
    static constexpr int JP = (1 << LS_LAST) + (1 << (LS_LAST - LOG2P));   // pad(b + (j << LS_LAST)) = pad(b) + JP*j
    static_assert(G::NPASS == 3 && T % 32 == 0 && T % P == 0, "geometry");
};

#if defined(__CUDACC__)

// 16-byte table entry that must not displace the per-pass twiddles from L1: the split tables (64 KB
// each at n = 16384) are read once per series per thread, the twiddles 62 times
__device__ __forceinline__ float4 load_f4_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void l2_prefetch(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Block-wide sums / maxima: every thread returns the same value (same order of operations).
template <int NWARP>
__device__ __forceinline__ float2 block_sum_f2(float a, float b, float2 *red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if ((tid & 31) == 0) red[tid >> 5] = make_float2(a, b);
    __syncthreads();
    float2 s = red[0];
#pragma unroll
    for (int w = 1; w < NWARP; w++) {
        s.x += red[w].x;
        s.y += red[w].y;
    }
    __syncthreads();
    return s;
}
template <int NWARP>
__device__ __forceinline__ float2 block_max_f2(float a, float b, float2 *red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, off));
        b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, off));
    }
    if ((tid & 31) == 0) red[tid >> 5] = make_float2(a, b);
    __syncthreads();
    float2 s = red[0];
#pragma unroll
    for (int w = 1; w < NWARP; w++) {
        s.x = fmaxf(s.x, red[w].x);
        s.y = fmaxf(s.y, red[w].y);
    }
    __syncthreads();
    return s;
}

// Forward FFT_M of the P points per thread in v (input j of thread t = element t + T*j); on return
// v[c*R_LAST + Perm<R_LAST>(j)] = Z[(t + c*T) + (j << LS_LAST)].  Ends WITHOUT a barrier after the last
// pass' loads: the caller synchronises before it writes the exchange buffer again.
//
// Same passes as fft_pass_compute_store / fft_pass_load of muse_fft.cuh for Geo<LOG2M, LOG2P>, with the
// padded shared-memory indices written out as (per-thread base) + (compile-time offset): T is a
// multiple of P, so pad(t + T*x) = pad(t) + TP*x, and pad(q + P*P*p + P*j) = q + (P*P + P)*p + (P + 1)*j.
// (The generic index arithmetic was 20 % of the kernel's instructions.)
template <int LOG2M>
__device__ __forceinline__ void block_fft(cf *v, cf *sm, int t, const cf *twp) {
    using C = ScreenBlockCfg<LOG2M>;
    using G = typename C::G;
    constexpr int T = C::T, TP = C::TP, P = C::P, LP = C::LOG2P;
    const int pt = t + (t >> LP);                           // pad(t)
    // pass 0: radix P over elements t + T*j, twiddle W_M^(j*t), scatter to P*t + j
    Dft<P, float>::run(v);
    {
        cf *dst = sm + (P + 1) * t;
        const cf *tw = twp + t;                             // row j-1, column t of the pass-0 table (T columns)
        dst[0] = v[Perm<P>::at(0)];
#pragma unroll
        for (int j = 1; j < P; j++) dst[j] = cmul(v[Perm<P>::at(j)], tw[(j - 1) * T]);
    }
    __syncthreads();
    // pass 1: butterfly b = t: p = t / P, q = t % P; inputs t + T*j; twiddle W_T^(j*p); scatter to q + P*P*p + P*j
#pragma unroll
    for (int j = 0; j < P; j++) v[j] = sm[pt + TP * j];
    __syncthreads();
    Dft<P, float>::run(v);
    {
        const int p = t >> LP, q = t & (P - 1);
        cf *dst = sm + q + (P * P + P) * p;
        const cf *tw = twp + G::tw_off(1) + p;              // row j-1, column p of the pass-1 table (T/P columns)
        dst[0] = v[Perm<P>::at(0)];
#pragma unroll
        for (int j = 1; j < P; j++) dst[(P + 1) * j] = cmul(v[Perm<P>::at(j)], tw[(j - 1) * (T / P)]);
    }
    __syncthreads();
    // last pass: NB_LAST butterflies of radix R_LAST per thread, b = t + c*T, inputs b + (j << LS_LAST); no twiddles
#pragma unroll
    for (int c = 0; c < C::NB_LAST; c++)
#pragma unroll
        for (int j = 0; j < C::R_LAST; j++) v[c * C::R_LAST + j] = sm[pt + TP * c + C::JP * j];
#pragma unroll
    for (int c = 0; c < C::NB_LAST; c++) Dft<C::R_LAST, float>::run(v + c * C::R_LAST);
}

template <int LOG2M, int MINB>
__global__ void __launch_bounds__(ScreenBlockCfg<LOG2M>::T, MINB)
score_screen_block_kernel(const ScreenParams prm) {
    using C = ScreenBlockCfg<LOG2M>;
    using G = typename C::G;
    constexpr int P = C::P, M = G::M, T = C::T, NW = C::NWARP, LP = C::LOG2P;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float2 red_f[NW];
    __shared__ unsigned bc_word[1];                  // running cut-off, broadcast by thread 0
    cf *sm = reinterpret_cast<cf *>(smem_raw);

    const int t = threadIdx.x;
    constexpr int TP = C::TP;                           // pad(t + T*x) = pad(t) + TP*x (T is a multiple of P)
    const int pt = t + (t >> LP);                       // pad(t)
    const int pm = M + (M >> LP) - t - ((t + P - 1) >> LP);   // pad(M - t) for t >= 1 (t = 0 pairs bin 0 with itself)
    const int N = prm.N;
    const int Nh = (N + 1) >> 1;      // complex slots holding samples (odd N: the pad column of the last one holds the row's mean, RowStat)
    const int count = (int)prm.count;
    const unsigned row_bytes = (unsigned)(N + (N & 1)) * 8u;

    for (int pos = blockIdx.x; pos < count; pos += gridDim.x) {
        const double *rowp = prm.slab + (int64_t)pos * prm.ld;
        unsigned cut_raw = 0u;
        if (t == 0) {
            if (pos + (int)gridDim.x < count) l2_prefetch(prm.slab + (int64_t)(pos + (int)gridDim.x) * prm.ld, row_bytes);
            cut_raw = ld_relaxed_u32(prm.cut_bits);
        }

        // ---- centred samples -> fp32 registers.  The fp64 mean comes from the ingest pass
        //      (row_stats_kernel), so there is no block reduction before the transform; loads go
        //      out in batches of 8 rows (8 x 16 bytes in flight per thread) ahead of their use ----
        const RowStat rs = prm.row_stat[pos];
        const double mu = rs.mean;
        cf v[P];
#pragma unroll
        for (int b0 = 0; b0 < P; b0 += 8) {              // P is 8, 16 or 32
            cd x[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = t + (b0 + q) * T;
                x[q] = j < Nh ? load_pair_stream(rowp + 2 * j) : cd{mu, mu};
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v[b0 + q] = cf{(float)(x[q].x - mu), (float)(x[q].y - mu)};      // exactly 0 in the padding
            }
        }

        // ---- forward FFT_M ----
        block_fft<LOG2M>(v, sm, t, prm.twp);
        __syncthreads();                                           // last pass' loads are done
#pragma unroll
        for (int c = 0; c < C::NB_LAST; c++)
#pragma unroll
            for (int j = 0; j < C::R_LAST; j++)
                sm[pt + TP * c + C::JP * j] = v[c * C::R_LAST + Perm<C::R_LAST>::at(j)];      // pad((t + c*T) + (j << LS_LAST))
        __syncthreads();

        // ---- |2Y_k| and |2Y_(M-k)| for k = t + T*i < M/2 (k = 0 pairs with itself: DC and Nyquist) ----
        // (the 16 table entries are fetched up front: consumed one by one in a rolled loop, each multiply
        //  waited for its own L2 round trip -- 9 % of all stall samples)
        float4 swr[P / 2];
#pragma unroll
        for (int i = 0; i < P / 2; i++) swr[i] = load_f4_stream(prm.sw + t + T * i);    // (w_k.x, w_k.y, A[k], A[M-k])
        cf acc2{0.f, 0.f};
#pragma unroll
        for (int i = 0; i < P / 2; i++) {
            const cf zk = sm[pt + TP * i];                             // pad(k)
            const cf zm = sm[(i == 0 && t == 0) ? 0 : pm - TP * i];    // pad(M - k)
            const float4 s = swr[i];
            const cf zmc = cconj(zm);
            const cf e = cadd(zk, zmc);
            const cf o = cmul_negi(csub(zk, zmc));
            const cf wo = cmul(o, cf{s.x, s.y});
            const cf y1 = cadd(e, wo);                             // 2*Y_k
            const cf y2 = csub(e, wo);                             // 2*conj(Y_(M-k))
            const cf q1 = pmul(y1, y1), q2 = pmul(y2, y2);
            const cf mag{sqrt_approx(q1.x + q1.y), sqrt_approx(q2.x + q2.y)};
            acc2 = pfma(mag, cf{s.z, s.w}, acc2);
        }
        float acc = acc2.x + acc2.y;
        if (t == 0) {                                              // k = M/2, its own mirror: |2Y| = 2|Z|
            const cf z = sm[G::pad(M / 2)];
            const cf q = pmul(z, z);
            acc = fmaf(sqrt_approx(q.x + q.y), 2.f * prm.a_mid, acc);
            bc_word[0] = cut_raw;
        }
        const float2 sums = block_sum_f2<NW>(acc, 0.f, red_f, t);   // barrier inside: bc_word is visible
        acc = sums.x;
        const float cut_now = __uint_as_float(bc_word[0]);
        // rstd is NaN for a row no fp32 statement may be made about (RowStat); NaN/Inf samples make acc NaN:
        // either way the bound is NaN and the exact kernel decides
        float U = acc * rs.rstd * 1.00001f + MUSE_SCREEN_SLACK;
        if (!(U == U)) U = 2.f;
        float L = -1.f;
        if (U >= cut_now && U < 1.5f) {                            // block-uniform
            // ---- conj(Y)*X on the mirror pairs, in place in shared memory (each pair has one owner) ----
            float4 sxr[P / 2];
#pragma unroll
            for (int i = 0; i < P / 2; i++) sxr[i] = load_f4_stream(prm.sx + t + T * i);
#pragma unroll
            for (int i = 0; i < P / 2; i++) {
                const int ik = pt + TP * i;                            // pad(k), k = t + T*i
                const int im = (i == 0 && t == 0) ? 0 : pm - TP * i;   // pad(M - k)
                const cf zk = sm[ik];
                const cf zm = sm[im];
                const float4 s = swr[i];
                const float4 x = sxr[i];
                cf ok, om;
                pointwise_pair(zk, zm, cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                sm[im] = om;
                sm[ik] = ok;                                           // k == M - k (k = 0): ok wins, as pointwise_phase
            }
            if (t == 0) {                                          // k = M/2; w = exp(-i*pi/2) = -i
                const cf mid = sm[G::pad(M / 2)];
                cf ok, om;
                pointwise_pair(mid, mid, cf{0.f, -1.f}, prm.x_mid, prm.x_mid, ok, om);
                sm[G::pad(M / 2)] = ok;
            }
            __syncthreads();
            // ---- inverse FFT_M as swap(FFT(swap(.))) ----
#pragma unroll
            for (int j = 0; j < P; j++) v[j] = sm[pt + TP * j];
            __syncthreads();
            block_fft<LOG2M>(v, sm, t, prm.twp);
            // v[c*R + Perm(j)] = (cc'[2k+1], cc'[2k]), k = (t + c*T) + (j << 10); cc' = std * cc rotated by pad
            float m_in = 0.f, m_out = 0.f;
            const int base = 2 * t - prm.win_lo;
#pragma unroll
            for (int c = 0; c < C::NB_LAST; c++)
#pragma unroll
                for (int j = 0; j < C::R_LAST; j++) {
                    const cf r = v[c * C::R_LAST + Perm<C::R_LAST>::at(j)];
                    const int off = 2 * (c * T + (j << C::LS_LAST));
                    const bool in0 = ((base + off) & (2 * M - 1)) <= prm.win_len;
                    const bool in1 = ((base + off + 1) & (2 * M - 1)) <= prm.win_len;
                    const float a0 = fabsf(r.y), a1 = fabsf(r.x);
                    m_in = fmaxf(m_in, fmaxf(in0 ? a0 : 0.f, in1 ? a1 : 0.f));
                    m_out = fmaxf(m_out, fmaxf(in0 ? 0.f : a0, in1 ? 0.f : a1));
                }
            const float2 mx = block_max_f2<NW>(m_in, m_out, red_f, t);
            const float rstd = rs.rstd;
            U = refine_decide(U, mx.x * rstd, mx.y * rstd, L, prm.grouped);
            if (t == 0) atomicAdd(prm.n_refined, 1ull);
            if (t < 32 && L >= prm.thr && L >= cut_now) cut_count_and_raise(prm, L, t);
        }
        if (t == 0) {
            prm.out_U[pos] = U;
            if (prm.out_L) prm.out_L[pos] = L;
        }
        __syncthreads();                                           // the exchange buffer and bc_word are free again
    }
}

#endif  // __CUDACC__

}  // namespace muse
