// muse_screen_wide.cuh -- the fp32 screening + fused second stage for FFT length 16384 (series of 8194 .. 16384 samples:
// BASELINE.json configs[3], one week at one sample per minute) at TWICE the occupancy of muse_screen_big.cuh.
//
// Same mathematics, contract and per-group running lower bound as score_screen_big_kernel.  That kernel keeps 32 points
// per thread (128 registers, 16 warps per SM) and measured at 38 % issue utilisation: with four warps per scheduler the
// latencies of the row loads, the table loads and the block barriers are not hidden (profiles/r02_big_*_ncu.txt).  Here a
// series is M = 8192 complex points = 16 per thread on 512 threads, 64 registers per thread, two blocks = 32 warps per SM:
// one more pass through shared memory (radix 16, 16, 16, 2 instead of 32, 32, 8: ~12 % more instructions), twice the
// warps to cover every wait.
//   * forward: Stockham radix 16, 16, 16, then a mirror-paired radix-2 pass: a thread transforms the butterflies b and
//     4096 - b, whose outputs k = b + 4096 j and M - k are the mirror pairs of the real split, so the split, the magnitudes
//     and (second stage) conj(Y) X run in registers;
//   * second stage: the inverse as a forward transform of the swapped values in the transposed pass order (2, 16, 16, 16),
//     straight from that register layout;
//   * every twiddle table is small enough to stay in L1 (the big passes' twiddles are factorised, W^(j t) = W^(4a t) W^(b t)).
// All per-thread phases are __host__ __device__ (tests/cpp/emulate_wide.cpp runs them thread by thread on the CPU).
#pragma once

#include "muse_screen_big.cuh"

namespace muse {

struct ScreenWideCfg {
    static constexpr int LOG2M = 13;
    static constexpr int M = 1 << LOG2M;              // 8192
    static constexpr int P = 16;
    static constexpr int T = M / P;                   // 512 threads per series
    static constexpr int NWARP = T / 32;
    static constexpr int NPAIR = 4;                   // mirror pairs of radix-2 butterflies per thread (u = t + 512 c2)
    static constexpr int TP = T + (T >> 4);           // pad(t + T x) = pad(t) + TP x,  pad(i) = i + (i >> 4)
    static constexpr int SM_ELEMS = M + (M >> 4) + 16;
    static constexpr size_t SMEM = (size_t)SM_ELEMS * 8;
    // fp32 twiddle tables, one array (fill_wide_twiddles)
    static constexpr int F0A = 0;                     // [3][512] W_M^(4 a t), a = 1..3
    static constexpr int F0B = F0A + 3 * T;           // [3][512] W_M^(b t), b = 1..3
    static constexpr int F1 = F0B + 3 * T;            // [15][32] W_512^(j p)
    static constexpr int F2 = F1 + 15 * 32;           // [15][2]  W_32^(j p)
    static constexpr int I0 = F2 + 15 * 2;            // [2048]   W_M^u
    static constexpr int I1A = I0 + 2048;             // [3][256] W_4096^(4 a p)
    static constexpr int I1B = I1A + 3 * 256;         // [3][256] W_4096^(b p)
    static constexpr int I2 = I1B + 3 * 256;          // [15][16] W_256^(j p)
    static constexpr int TW_TOTAL = I2 + 15 * 16;
};

template <typename TW, typename FN>
inline void fill_wide_twiddles(TW *out, FN unit_root /* (num, den) -> TW */) {
    using C = ScreenWideCfg;
    for (int a = 1; a < 4; a++)
        for (int t = 0; t < C::T; t++) {
            out[C::F0A + (a - 1) * C::T + t] = unit_root((long long)4 * a * t, (long long)C::M);
            out[C::F0B + (a - 1) * C::T + t] = unit_root((long long)a * t, (long long)C::M);
        }
    for (int j = 1; j < 16; j++) {
        for (int p = 0; p < 32; p++) out[C::F1 + (j - 1) * 32 + p] = unit_root((long long)j * p, 512LL);
        for (int p = 0; p < 2; p++) out[C::F2 + (j - 1) * 2 + p] = unit_root((long long)j * p, 32LL);
        for (int p = 0; p < 16; p++) out[C::I2 + (j - 1) * 16 + p] = unit_root((long long)j * p, 256LL);
    }
    for (int u = 0; u < 2048; u++) out[C::I0 + u] = unit_root((long long)u, (long long)C::M);
    for (int a = 1; a < 4; a++)
        for (int p = 0; p < 256; p++) {
            out[C::I1A + (a - 1) * 256 + p] = unit_root((long long)4 * a * p, 4096LL);
            out[C::I1B + (a - 1) * 256 + p] = unit_root((long long)a * p, 4096LL);
        }
}

MUSE_HD int wide_pad(int i) { return i + (i >> 4); }

// v[Perm<16>(j)] *= W^(j x) for j = 1..15 with W^(j x) = A[a] B[b], j = 4 a + b; tabA / tabB point at column x (row stride `stride`)
MUSE_HD void wide_twiddle_factored(cf *v, const cf *tabA, const cf *tabB, int stride) {
    cf wb[4];
#pragma unroll
    for (int b = 1; b < 4; b++) {
        wb[b] = tabB[(b - 1) * stride];
        v[Perm<16>::at(b)] = cmul(v[Perm<16>::at(b)], wb[b]);
    }
#pragma unroll
    for (int a = 1; a < 4; a++) {
        const cf wa = tabA[(a - 1) * stride];
        v[Perm<16>::at(4 * a)] = cmul(v[Perm<16>::at(4 * a)], wa);
#pragma unroll
        for (int b = 1; b < 4; b++) v[Perm<16>::at(4 * a + b)] = cmul(v[Perm<16>::at(4 * a + b)], cmul(wa, wb[b]));
    }
}

// inputs of a pass whose butterfly is the thread itself: elements t + 512 j under pad
MUSE_HD void wide_load_stride_t(cf *v, const cf *sm, int t) {
    const int pt = wide_pad(t);
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = sm[pt + ScreenWideCfg::TP * j];
}

// ---- forward ----------------------------------------------------------------------------------------------------------
// pass 0: inputs v[j] = z[t + 512 j]; writes y[16 t + j] W_M^(j t)
MUSE_HD void wide_fwd_pass0(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
    Dft<16, float>::run(v);
    wide_twiddle_factored(v, tw + C::F0A + t, tw + C::F0B + t, C::T);
    cf *dst = sm + 17 * t;
#pragma unroll
    for (int j = 0; j < 16; j++) dst[j] = v[Perm<16>::at(j)];
}
// pass 1: p = t / 16, q = t % 16; writes y[q + 256 p + 16 j] W_512^(j p)
MUSE_HD void wide_fwd_pass1(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
    Dft<16, float>::run(v);
    const int p = t >> 4, q = t & 15;
    cf *dst = sm + q + 272 * p;
    const cf *w = tw + C::F1 + p;
    dst[0] = v[Perm<16>::at(0)];
#pragma unroll
    for (int j = 1; j < 16; j++) dst[17 * j] = cmul(v[Perm<16>::at(j)], w[(j - 1) * 32]);
}
// pass 2: p = t / 256, q = t % 256; writes y[q + 4096 p + 256 j] W_32^(j p)
MUSE_HD void wide_fwd_pass2(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
    Dft<16, float>::run(v);
    const int p = t >> 8, q = t & 255;
    cf *dst = sm + q + (q >> 4) + 4352 * p;
    const cf *w = tw + C::F2 + p;
    dst[0] = v[Perm<16>::at(0)];
#pragma unroll
    for (int j = 1; j < 16; j++) dst[272 * j] = cmul(v[Perm<16>::at(j)], w[(j - 1) * 2]);
}
// butterflies of the last (radix 2) pass that thread t owns: pair slot c2 holds b_lo = u and b_hi = 4096 - u, u = t + 512 c2
// (u == 0: b_lo = 0 and b_hi = 2048, the two butterflies that are their own mirrors)
MUSE_HD int wide_b_hi(int u) { return u == 0 ? 2048 : 4096 - u; }
// last pass: on return v[(2 c2 + h) 2 + j] = Z[b + 4096 j]
MUSE_HD void wide_fwd_last(cf *v, const cf *sm, int t) {
    using C = ScreenWideCfg;
#pragma unroll
    for (int c2 = 0; c2 < C::NPAIR; c2++) {
        const int u = t + C::T * c2;
        const int bh = wide_b_hi(u);
        const int plo = wide_pad(u), phi = wide_pad(bh);
        const cf a0 = sm[plo], a1 = sm[plo + 4352], b0 = sm[phi], b1 = sm[phi + 4352];
        v[4 * c2 + 0] = cadd(a0, a1);
        v[4 * c2 + 1] = csub(a0, a1);
        v[4 * c2 + 2] = cadd(b0, b1);
        v[4 * c2 + 3] = csub(b0, b1);
    }
}

// the thread's share of sum_k |2Y_k| A_k.  Pair slot c2: k = u (lo[0]) pairs with M - u (hi[1]), k' = 4096 - u (hi[0]) with
// M - k' = u + 4096 (lo[1]).  u == 0: lo[0] = Z[0] pairs with itself (DC and Nyquist), lo[1] = Z[M/2] is its own mirror
// (|2Y| = 2|Z|), hi = (Z[2048], Z[6144]) is one pair.
MUSE_HD float wide_split_bound(const cf *v, int t, const float4 *sw, float a_mid) {
    using C = ScreenWideCfg;
    cf acc2{0.f, 0.f};
    float extra = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < C::NPAIR; c2++) {
        const int u = t + C::T * c2;
        const cf *lo = v + 4 * c2, *hi = v + 4 * c2 + 2;
        if (c2 == 0 && u == 0) {
            big_split_acc(lo[0], lo[0], big_load_f4(sw), acc2);
            const cf q = pmul(lo[1], lo[1]);
            extra = big_sqrt(q.x + q.y) * (2.f * a_mid);
            big_split_acc(hi[0], hi[1], big_load_f4(sw + 2048), acc2);
        } else {
            big_split_acc(lo[0], hi[1], big_load_f4(sw + u), acc2);
            big_split_acc(hi[0], lo[1], big_load_f4(sw + (4096 - u)), acc2);
        }
    }
    return acc2.x + acc2.y + extra;
}

// second stage, in place: conj(Y) X on every mirror pair (pointwise_pair stores the swapped values)
MUSE_HD void wide_pointwise(cf *v, int t, const float4 *sw, const float4 *sx, cf x_mid) {
    using C = ScreenWideCfg;
#pragma unroll
    for (int c2 = 0; c2 < C::NPAIR; c2++) {
        const int u = t + C::T * c2;
        cf *lo = v + 4 * c2, *hi = v + 4 * c2 + 2;
        cf ok, om;
        if (c2 == 0 && u == 0) {
            {
                const float4 s = big_load_f4(sw), x = big_load_f4(sx);
                pointwise_pair(lo[0], lo[0], cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                lo[0] = ok;
            }
            pointwise_pair(lo[1], lo[1], cf{0.f, -1.f}, x_mid, x_mid, ok, om);      // bin M/2: w = -i
            lo[1] = ok;
            {
                const float4 s = big_load_f4(sw + 2048), x = big_load_f4(sx + 2048);
                pointwise_pair(hi[0], hi[1], cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                hi[0] = ok;
                hi[1] = om;
            }
        } else {
            {
                const float4 s = big_load_f4(sw + u), x = big_load_f4(sx + u);
                pointwise_pair(lo[0], hi[1], cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                lo[0] = ok;
                hi[1] = om;
            }
            {
                const float4 s = big_load_f4(sw + (4096 - u)), x = big_load_f4(sx + (4096 - u));
                pointwise_pair(hi[0], lo[1], cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                hi[0] = ok;
                lo[1] = om;
            }
        }
    }
}

// ---- inverse as a forward transform of the swapped values, pass order 2, 16, 16, 16 -----------------------------------------
// pass 0': butterfly b holds x[b], x[b + 4096]; writes y[2 b], y[2 b + 1] W_M^b under pad2(i) = i + 2 (i >> 5)
// (W_M^(4096 - u) = -conj(W_M^u), W_M^2048 = -i)
MUSE_HD void wide_inv_pass0(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
#pragma unroll
    for (int c2 = 0; c2 < C::NPAIR; c2++) {
        const int u = t + C::T * c2;
        const int bh = wide_b_hi(u);
        const cf wu = tw[C::I0 + u];
        const cf wh = (u == 0) ? cf{0.f, -1.f} : cf{-wu.x, wu.y};
        {
            const cf a = v[4 * c2], b = v[4 * c2 + 1];
            cf *dst = sm + 2 * u + 2 * (u >> 4);
            dst[0] = cadd(a, b);
            dst[1] = cmul(csub(a, b), wu);
        }
        {
            const cf a = v[4 * c2 + 2], b = v[4 * c2 + 3];
            cf *dst = sm + 2 * bh + 2 * (bh >> 4);
            dst[0] = cadd(a, b);
            dst[1] = cmul(csub(a, b), wh);
        }
    }
}
MUSE_HD void wide_inv_pass1_load(cf *v, const cf *sm, int t) {
    const int pt2 = t + 2 * (t >> 5);
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = sm[pt2 + ScreenWideCfg::TP * j];
}
// pass 1': p = t / 2, q = t % 2; writes y[q + 32 p + 2 j] W_4096^(j p) under pad
MUSE_HD void wide_inv_pass1(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
    Dft<16, float>::run(v);
    const int p = t >> 1, q = t & 1;
    wide_twiddle_factored(v, tw + C::I1A + p, tw + C::I1B + p, 256);
    cf *dst = sm + q + 34 * p;
#pragma unroll
    for (int j = 0; j < 16; j++) dst[2 * j + (j >= 8 ? 1 : 0)] = v[Perm<16>::at(j)];
}
// pass 2': p = t / 32, q = t % 32; writes y[q + 512 p + 32 j] W_256^(j p) under pad
MUSE_HD void wide_inv_pass2(cf *v, cf *sm, int t, const cf *tw) {
    using C = ScreenWideCfg;
    Dft<16, float>::run(v);
    const int p = t >> 5, q = t & 31;
    cf *dst = sm + q + (q >> 4) + 544 * p;
    const cf *w = tw + C::I2 + p;
    dst[0] = v[Perm<16>::at(0)];
#pragma unroll
    for (int j = 1; j < 16; j++) dst[34 * j] = cmul(v[Perm<16>::at(j)], w[(j - 1) * 16]);
}
// pass 3' (after wide_load_stride_t): radix 16, no twiddles: v[Perm<16>(j)] = (cc'[2i+1], cc'[2i]), i = t + 512 j
MUSE_HD void wide_window_max(const cf *v, int t, int win_lo, int win_len, float &m_in, float &m_out) {
    using C = ScreenWideCfg;
    m_in = 0.f;
    m_out = 0.f;
    const int base = 2 * t - win_lo;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const cf r = v[Perm<16>::at(j)];
        const int off = 2 * C::T * j;
        const bool in0 = ((base + off) & (2 * C::M - 1)) <= win_len;
        const bool in1 = ((base + off + 1) & (2 * C::M - 1)) <= win_len;
        const float a0 = fabsf(r.y), a1 = fabsf(r.x);
        m_in = fmaxf(m_in, fmaxf(in0 ? a0 : 0.f, in1 ? a1 : 0.f));
        m_out = fmaxf(m_out, fmaxf(in0 ? 0.f : a0, in1 ? 0.f : a1));
    }
}

#if defined(__CUDACC__) && defined(MUSE_WIDE_KERNEL)      // the kernel lives in ONE translation unit (kernels_screen_big.cu)

__global__ void __launch_bounds__(ScreenWideCfg::T, 2)
score_screen_wide_kernel(const ScreenParams prm) {
    using C = ScreenWideCfg;
    constexpr int P = C::P, T = C::T, NW = C::NWARP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float2 red_f[2][NW];
    __shared__ unsigned bc_word[2];
    cf *sm = reinterpret_cast<cf *>(smem_raw);

    const int t = threadIdx.x;
    const int N = prm.N;
    const int Nh = (N + 1) >> 1;      // complex slots holding samples (odd N: the pad column of the last one holds the row's mean, RowStat)
    const int count = (int)prm.count;
    const unsigned row_bytes = (unsigned)(N + (N & 1)) * 8u;
    const int chunk = (count + (int)gridDim.x - 1) / (int)gridDim.x;      // contiguous range of series per block
    const int pos_lo = (int)blockIdx.x * chunk;
    const int pos_hi = min(pos_lo + chunk, count);
    if (t == 0 && pos_lo < pos_hi) big_l2_prefetch(prm.slab + (int64_t)pos_lo * prm.ld, row_bytes);

    for (int pos = pos_lo; pos < pos_hi; pos++) {
        const double *rowp = prm.slab + (int64_t)pos * prm.ld;
        unsigned cut_raw = 0u, lg_raw = 0u;
        unsigned long long *gslot = nullptr;
        if (t == 0) {
            if (pos + 1 < pos_hi) big_l2_prefetch(rowp + prm.ld, row_bytes);
            cut_raw = ld_relaxed_u32(prm.cut_bits);
            if (prm.group_L) {
                gslot = prm.group_L + prm.slot_of[pos];
                lg_raw = ld_relaxed_u32(reinterpret_cast<const unsigned *>(gslot));
            }
        }
        const RowStat rs = prm.row_stat[pos];
        const double mu = rs.mean;

        // ---- centred samples -> fp32 registers, 8 x 16 bytes in flight per thread ----
        cf v[P];
#pragma unroll
        for (int b0 = 0; b0 < P; b0 += 8) {
            cd x[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = t + (b0 + q) * T;
                x[q] = j < Nh ? load_pair_stream(rowp + 2 * j) : cd{mu, mu};
            }
#pragma unroll
            for (int q = 0; q < 8; q++) v[b0 + q] = cf{(float)(x[q].x - mu), (float)(x[q].y - mu)};
        }

        // ---- forward FFT_M; ends with Z in registers in mirror-paired order ----
        wide_fwd_pass0(v, sm, t, prm.twi);
        __syncthreads();
        wide_load_stride_t(v, sm, t);
        __syncthreads();
        wide_fwd_pass1(v, sm, t, prm.twi);
        __syncthreads();
        wide_load_stride_t(v, sm, t);
        __syncthreads();
        wide_fwd_pass2(v, sm, t, prm.twi);
        __syncthreads();
        wide_fwd_last(v, sm, t);

        // ---- bound ----
        float acc = wide_split_bound(v, t, prm.sw, prm.a_mid);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if ((t & 31) == 0) red_f[0][t >> 5].x = acc;
        if (t == 0) {
            bc_word[0] = cut_raw;
            bc_word[1] = lg_raw;
        }
        __syncthreads();                                             // also: every last-pass load of the exchange buffer is done
        acc = red_f[0][0].x;
#pragma unroll
        for (int w = 1; w < NW; w++) acc += red_f[0][w].x;
        const float cut_now = __uint_as_float(bc_word[0]);
        const float lg_now = __uint_as_float(bc_word[1]);
        float U = acc * rs.rstd * 1.00001f + MUSE_SCREEN_SLACK;
        if (!(U == U)) U = 2.f;
        float L = -1.f;
        signed char W = 0;
        if (U >= cut_now && U >= lg_now && U < 1.5f) {               // block-uniform
            wide_pointwise(v, t, prm.sw, prm.sx, prm.x_mid);
            wide_inv_pass0(v, sm, t, prm.twi);
            __syncthreads();
            wide_inv_pass1_load(v, sm, t);
            __syncthreads();
            wide_inv_pass1(v, sm, t, prm.twi);
            __syncthreads();
            wide_load_stride_t(v, sm, t);
            __syncthreads();
            wide_inv_pass2(v, sm, t, prm.twi);
            __syncthreads();
            wide_load_stride_t(v, sm, t);
            Dft<16, float>::run(v);
            float m_in, m_out;
            wide_window_max(v, t, prm.win_lo, prm.win_len, m_in, m_out);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, off));
                m_out = fmaxf(m_out, __shfl_xor_sync(0xffffffffu, m_out, off));
            }
            if ((t & 31) == 0) red_f[1][t >> 5] = make_float2(m_in, m_out);
            __syncthreads();
            float2 mx = red_f[1][0];
#pragma unroll
            for (int w = 1; w < NW; w++) {
                mx.x = fmaxf(mx.x, red_f[1][w].x);
                mx.y = fmaxf(mx.y, red_f[1][w].y);
            }
            const float rstd = rs.rstd;
            const float s_in = mx.x * rstd, s_out = mx.y * rstd;
            U = refine_decide(U, s_in, s_out, L, prm.grouped);
            if (prm.grouped) {
                float Lw = -1.f;
                const float uw = refine_decide(2.f, s_in, s_out, Lw, 0);
                W = uw < 0.f ? -1 : (Lw >= 0.f ? 1 : 0);
            }
            if (t == 0) {
                atomicAdd(prm.n_refined, 1ull);
                if (gslot && L > lg_now) atomicMax(gslot, (unsigned long long)__float_as_uint(L));
            }
            if (!prm.grouped && t < 32 && L >= prm.thr && L >= cut_now) cut_count_and_raise(prm, L, t);
        }
        if (t == 0) {
            prm.out_U[pos] = U;
            if (prm.out_L) prm.out_L[pos] = L;
            if (prm.out_W) prm.out_W[pos] = W;
        }
        // (as in score_screen_big_kernel: the exchange buffer, red_f and bc_word are not written again before later barriers)
    }
}

#endif  // __CUDACC__

}  // namespace muse
