// muse_screen_multi.cuh -- the fused fp32 screening pass of muse_screen.cuh for SEVERAL reference queries
// in ONE pass over the slab (SURVEY 8f rank 2 / BASELINE.json configs[4]: many references, one store).
//
// go-muse builds one Batch per reference (muse_batch.go:23-52), so Q queries stream the store Q times and
// transform every series Q times.  Nothing about a series' spectrum depends on the reference: here a warp
// loads its row and runs the forward FFT_1024 ONCE, keeps the 1024 magnitudes |2Y_k| in registers (as the 16
// mirror pairs + the middle bin of muse_screen.cuh) and the spectrum itself in shared memory (the warp's row buffer),
// and then walks the queries of the launch:
//     U_q = (1/n) sum_f |Y_f| |X_q,f| / std * (1 + 1e-5) + slack      (16 packed FMAs against the query's weights,
//                                                                       which sit in shared memory for the whole launch)
// and, when U_q reaches query q's RUNNING top-N cut-off, the second stage of muse_screen.cuh for that query
// (spectrum reloaded from the stash, conj(Y)*X_q, inverse FFT_1024, maxima inside / outside the lag window).
// Every query has its own cut-off state (muse_batch::d_cut) and its own bound array (muse_batch::d_U): after the
// launch each query's batch is exactly where a single-query run is after ITS screening kernel, and the
// unchanged tail (survivors -> exact fp64 kernel -> filter -> top-N) finishes it.  The error budget of the
// bounds is that of muse_screen.cuh (the same operations on the same values; only the order of the 33-term
// accumulation differs, which the 1e-5 relative slack covers 30 times over).
//
// Cost per (series, query): 8 LDS.128 + 16 FFMA2 + a 5-step butterfly ~ 50 warp instructions (four queries at a
// time, so that their dependency chains overlap), against ~1000 for the row's load + transform, which is paid
// once per launch instead of once per query.
#pragma once

#include "muse_screen.cuh"

namespace muse {

struct MultiQuery {
    const float4 *sw;     // the query's (w_k, A[k], A[M-k]) table (only the weights are read here)
    const float4 *sx;     // (Xt[k], Xt[M-k]) in fp32, k < M/2
    cf x_mid;             // Xt[M/2]
    float a_mid;          // A[M/2]
    int pad;
    unsigned *cut;        // the query's cut-off state: [0] running cut-off bits, [2..3] refined count, [4..] histogram
    float *out_U;         // [count] upper bound on the score for this query
};

#ifndef MUSE_MULTI_WARPS
#define MUSE_MULTI_WARPS 10   // measured (16 queries x 1 M series): 7 warps + separate stash 8.04 ms, 12 warps 7.93 ms, 10 warps 7.32 ms
#endif
struct ScreenMultiCfg {
    using G = Geo<10, 5>;
    static constexpr int MAX_WARPS = MUSE_MULTI_WARPS;   // 168 registers per thread, no spills
    static constexpr int QC = 16;               // queries per launch (their weights: 64 KB of shared memory)
    static constexpr int NREF = ScreenWarpCfg::NREF;
    static constexpr size_t SMEM_BUDGET = 227 * 1024;
    static constexpr size_t EX_BYTES = ScreenWarpCfg::EX_BYTES;
    static constexpr size_t STASH_BYTES = (size_t)G::M * sizeof(cf);
    static constexpr size_t W_BYTES = (size_t)QC * (G::M / 2) * sizeof(cf) + 128;      // weights + A[M/2] per query
    // one buffer per warp, three uses in turn: the row (bulk copy), the forward FFT exchange, the spectrum stash
    static size_t row_bytes(int N) { const size_t r = ScreenWarpCfg::row_bytes(N); return r > STASH_BYTES ? r : STASH_BYTES; }
    static size_t warp_bytes(int N) { return row_bytes(N); }
    static int warps(int N) {
        const size_t w = (SMEM_BUDGET - NREF * EX_BYTES - W_BYTES) / warp_bytes(N);
        return (int)(w > MAX_WARPS ? MAX_WARPS : w);
    }
    static size_t smem_bytes(int N) { return (size_t)warps(N) * warp_bytes(N) + NREF * EX_BYTES + W_BYTES; }
};

#if defined(__CUDACC__)

// cut_count_and_raise of muse_screen.cuh on an explicit cut-off state
__device__ __forceinline__ void cut_count_and_raise_at(unsigned *cut, int top_n, float L, int t) {
    ScreenParams q;
    q.cut_bits = cut;
    q.cut_hist = cut + 4;
    q.top_n = top_n;
    cut_count_and_raise(q, L, t);
}

template <int NZ>
__global__ void __launch_bounds__(ScreenMultiCfg::MAX_WARPS * 32, 1)
score_screen_multi_kernel(const ScreenParams prm, const MultiQuery *__restrict__ queries, const int nq, const unsigned warp_bytes,
                          const unsigned row_bytes) {
    using C = ScreenMultiCfg;
    using G = typename C::G;
    constexpr int P = 32, M = G::M;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[C::MAX_WARPS];
    __shared__ unsigned ex_locks[C::NREF];

    const int w = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    const int nwarps = (int)(blockDim.x >> 5);
    const int count = (int)prm.count;
    const int stride = (int)(gridDim.x * nwarps);
    const int pos0 = (int)(blockIdx.x * nwarps) + w;
    unsigned char *buf = smem_raw + (size_t)w * warp_bytes;
    const cd *rowc = reinterpret_cast<const cd *>(buf);
    cf *sm = reinterpret_cast<cf *>(buf);                       // forward exchange: the row buffer itself
    cf *stash = reinterpret_cast<cf *>(buf);                    // ... and then the series' spectrum, slot j of lane t at [32 j + t]
    unsigned char *refbase = smem_raw + (size_t)nwarps * warp_bytes;
    cf *sA = reinterpret_cast<cf *>(refbase + (size_t)C::NREF * C::EX_BYTES);     // [nq][512] (A[k], A[M-k])
    float *sAmid = reinterpret_cast<float *>(sA + (size_t)C::QC * (M / 2));       // [nq] A[M/2]
    const int N = prm.N;
    const int Nh = (N + 1) >> 1;      // complex slots holding samples (odd N: the pad column of the last one holds the row's mean, RowStat)
    const unsigned bar = smem_u32(&bars[w]);
    const int partner = (P - t) & (P - 1);
    const bool lane0 = (t == 0);
    const bool last_in = t + (NZ - 1) * 32 < Nh;

    if (threadIdx.x < C::NREF) ex_locks[threadIdx.x] = 0u;
    if (t == 0) {
        mbar_init(bar, 1);
        if (pos0 < count) bulk_load(smem_u32(buf), prm.slab + (int64_t)pos0 * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
    }
    // weights of query q for the mirror pairs k = t + 32 j: the pairs of slots j = 2 j2 and 2 j2 + 1 share one 16-byte
    // entry [q][j2][t] (one LDS.128 per two packed FMAs)
    for (int i = threadIdx.x; i < nq * (M / 2); i += blockDim.x) {
        const int q = i >> 9, k = i & (M / 2 - 1), j = k >> 5, tt = k & 31;
        const float4 s = queries[q].sw[k];
        sA[((q * (P / 4) + (j >> 1)) * 32 + tt) * 2 + (j & 1)] = cf{s.z, s.w};
    }
    if ((int)threadIdx.x < nq) sAmid[threadIdx.x] = queries[threadIdx.x].a_mid;
    // lane q looks after query q: its cut-off word and its row of bounds
    unsigned *my_cut = t < nq ? queries[t].cut : nullptr;
    float *my_out = t < nq ? queries[t].out_U : nullptr;
    __syncthreads();

    for (unsigned pp = (unsigned)pos0; (int)(pp & 0x7fffffffu) < count; pp = ((pp & 0x7fffffffu) + (unsigned)stride) | (~pp & 0x80000000u)) {
        const int pos = (int)(pp & 0x7fffffffu);
        const unsigned phase = pp >> 31;
        // the running cut-offs of all queries, read at the top of the iteration and consumed after the transform
        unsigned cut_raw = 0x7f800000u;     // +inf: no such query
        if (my_cut) cut_raw = ld_relaxed_u32(my_cut);
        const RowStat rs = prm.row_stat[pos];
        const double mu = rs.mean;
        mbar_wait(bar, phase);

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
        __syncwarp();
        const int next = pos + stride;

        // ---- forward FFT_1024 (as score_screen_warp_kernel) ----
        Dft32Lead<NZ, float>::run(v);
#pragma unroll
        for (int j = 0; j < P; j++) {
            cf val = v[Perm<P>::at(j)];
            if (j > 0) val = cmul(val, prm.twp[(j - 1) * 32 + t]);
            sm[G::pad(32 * t + j)] = val;
        }
        __syncwarp();
        fft_pass_load<10, 5, 1, float>(v, sm, t);
        __syncwarp();
        Dft<P, float>::run(v);                          // v[Perm(j)] = Z[t + 32*j]

        // ---- the spectrum goes to the stash: the same buffer once more (the exchange has been read, __syncwarp above;
        //      every lane reads back only what it wrote) ----
#pragma unroll
        for (int j = 0; j < P; j++) stash[32 * j + t] = v[Perm<P>::at(j)];

        // ---- magnitudes (|2Y_k|, |2Y_(M-k)|), k = t + 32*j, j < 16, and |Z_512| ----
        cf mg[P / 2];
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
            const cf zmc = cconj(zm);
            const cf e = cadd(zk, zmc);
            const cf o = cmul_negi(csub(zk, zmc));
            const cf wo = cmul(o, cf{s.x, s.y});
            const cf y1 = cadd(e, wo);
            const cf y2 = csub(e, wo);
            const cf q1 = pmul(y1, y1), q2 = pmul(y2, y2);
            mg[j] = cf{sqrt_approx(q1.x + q1.y), sqrt_approx(q2.x + q2.y)};
        }
        float mg_mid;
        {
            const cf z = v[Perm<P>::at(P / 2)];
            const cf q = pmul(z, z);
            mg_mid = lane0 ? 2.f * sqrt_approx(q.x + q.y) : 0.f;      // k = 512: lane 0, slot 16
        }

        // ---- all bounds first (lane q keeps query q's), four queries per step: their 8 accumulation chains and 4
        //      butterfly reductions overlap (one at a time leaves a warp waiting on 8 FFMA2 and 5 dependent shuffles) ----
        float my_U = 2.f;
        for (int q = 0; q < nq; q += 4) {
            float acc4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int qi = q + i < nq ? q + i : nq - 1;
                const float4 *Aq = reinterpret_cast<const float4 *>(sA) + (size_t)qi * (M / 4) + t;
                cf a0{0.f, 0.f}, a1{0.f, 0.f};
#pragma unroll
                for (int j2 = 0; j2 < P / 4; j2++) {
                    const float4 wq = Aq[32 * j2];
                    a0 = pfma(mg[2 * j2], cf{wq.x, wq.y}, a0);
                    a1 = pfma(mg[2 * j2 + 1], cf{wq.z, wq.w}, a1);
                }
                acc4[i] = fmaf(mg_mid, sAmid[qi], (a0.x + a0.y) + (a1.x + a1.y));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int i = 0; i < 4; i++) acc4[i] += __shfl_xor_sync(0xffffffffu, acc4[i], off);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float U = acc4[i] * rs.rstd * 1.00001f + MUSE_SCREEN_SLACK;
                if (!(U == U)) U = 2.f;
                if (t == q + i) my_U = U;      // q + i >= nq: a duplicate of the last query on a lane without one
            }
        }
        // ---- then the second stage for every query whose bound reaches ITS running cut-off (the magnitudes are dead
        //      by now; consecutive refinements of one series run the same code back to back) ----
        unsigned todo = __ballot_sync(0xffffffffu, t < nq && my_U >= __uint_as_float(cut_raw) && my_U < 1.5f);
        while (todo) {
            const int q = __ffs(todo) - 1;
            todo &= todo - 1;
            float U = __shfl_sync(0xffffffffu, my_U, q);
            const float cut_now = __uint_as_float(__shfl_sync(0xffffffffu, cut_raw, q));
            {
                // ---- second stage for query q: spectrum back from the stash, conj(Y)*X_q, inverse FFT, window maxima ----
                const MultiQuery mq = queries[q];
                cf z[P];
#pragma unroll
                for (int j = 0; j < P; j++) z[Perm<P>::at(j)] = stash[32 * j + t];
#pragma unroll
                for (int j = 0; j < P / 2; j++) {
                    const cf zk = z[Perm<P>::at(j)];
                    const cf zp = z[Perm<P>::at(P - 1 - j)];
                    const cf zs = z[Perm<P>::at((P - j) & (P - 1))];
                    cf src, zm;
                    src.x = lane0 ? zs.x : zp.x;
                    src.y = lane0 ? zs.y : zp.y;
                    zm.x = __shfl_sync(0xffffffffu, src.x, partner);
                    zm.y = __shfl_sync(0xffffffffu, src.y, partner);
                    const float4 s = prm.sw[t + 32 * j];
                    const float4 x = mq.sx[t + 32 * j];
                    cf ok, om;
                    pointwise_pair(zk, zm, cf{s.x, s.y}, cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                    cf rcv;
                    rcv.x = __shfl_sync(0xffffffffu, om.x, partner);
                    rcv.y = __shfl_sync(0xffffffffu, om.y, partner);
                    z[Perm<P>::at(j)] = ok;
                    cf &hi = z[Perm<P>::at(P - 1 - j)];
                    hi.x = lane0 ? hi.x : rcv.x;
                    hi.y = lane0 ? hi.y : rcv.y;
                    if (j > 0) {
                        cf &own = z[Perm<P>::at(P - j)];
                        own.x = lane0 ? om.x : own.x;
                        own.y = lane0 ? om.y : own.y;
                    }
                }
                {
                    cf &mid = z[Perm<P>::at(P / 2)];
                    cf ok, om;
                    pointwise_pair(mid, mid, cf{0.f, -1.f}, mq.x_mid, mq.x_mid, ok, om);
                    mid.x = lane0 ? ok.x : mid.x;
                    mid.y = lane0 ? ok.y : mid.y;
                }
                cf u[P];
#pragma unroll
                for (int j = 0; j < P; j++) u[j] = z[Perm<P>::at(j)];
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
                Dft<P, float>::run(u);
                float m_in = 0.f, m_out = 0.f;
                const int base = 2 * t - prm.win_lo;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const cf r = u[Perm<P>::at(j)];
                    const bool in0 = ((base + 64 * j) & (2 * M - 1)) <= prm.win_len;
                    const bool in1 = ((base + 64 * j + 1) & (2 * M - 1)) <= prm.win_len;
                    const float b0 = fabsf(r.y), b1 = fabsf(r.x);
                    m_in = fmaxf(m_in, in0 ? b0 : 0.f);
                    m_out = fmaxf(m_out, in0 ? 0.f : b0);
                    m_in = fmaxf(m_in, in1 ? b1 : 0.f);
                    m_out = fmaxf(m_out, in1 ? 0.f : b1);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, off));
                    m_out = fmaxf(m_out, __shfl_xor_sync(0xffffffffu, m_out, off));
                }
                float L = -1.f;
                U = refine_decide(U, m_in * rs.rstd, m_out * rs.rstd, L, 0);
                if (t == 0) atomicAdd(reinterpret_cast<unsigned long long *>(mq.cut + 2), 1ull);
                if (L >= prm.thr && L >= cut_now) cut_count_and_raise_at(mq.cut, prm.top_n, L, t);
            }
            if (t == q) my_U = U;
        }
        if (my_out) my_out[pos] = my_U;
        // the stash is no longer needed: hand the buffer to the copy engine for the warp's next row.  (No prefetch under
        // the transform as in 5.2: a separate stash costs 8 KB per warp, i.e. 7 warps per SM instead of 10 - 12, and this
        // kernel is bound by latency hiding, not by DRAM.)
        __syncwarp();
        if (t == 0 && next < count) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(smem_u32(buf), prm.slab + (int64_t)next * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
        }
    }
}


// ---- second stage for up to 256 queries whose bounds were computed elsewhere (muse_bounds_tc.cuh) -------------------
// One pass over the store: a warp loads its row, runs the forward FFT_1024 once, stashes the spectrum in shared memory
// and reads the series' bound against every query of the launch (query q's array out_U[pos], lane l looks after the
// queries l, l + 32, ...).  Every (series, query) pair whose bound reaches that query's RUNNING cut-off takes the second
// stage of muse_screen.cuh (conj(Y) X_q, inverse FFT_1024, maxima inside / outside the lag window), which replaces the
// bound by the tight fp32 one and feeds the query's cut-off histogram.  Afterwards every query is exactly where a
// single-query run is after its screening kernel.  Each warp owns its exchange buffer: with ~10 refinements per series a
// shared, locked buffer would serialise the block.
struct RefineMultiCfg {
    using G = Geo<10, 5>;
#ifndef MUSE_REFINE_WARPS
#define MUSE_REFINE_WARPS 8      // measured (256 queries x 1 M series): 12 warps at 168 registers 112 ms, 10 warps 76 ms, 8 warps at 255 registers 72 ms
#endif
    static constexpr int MAX_WARPS = MUSE_REFINE_WARPS;
    static constexpr int QMAX = 256;            // queries per launch (== TcCfg::TN)
    static constexpr int QI = QMAX / 32;        // bounds per lane
    static constexpr size_t SMEM_BUDGET = 227 * 1024;
    static constexpr size_t STASH_BYTES = (size_t)G::M * sizeof(cf);                                  // 8 KB: the spectrum
    static constexpr size_t EX_BYTES = ScreenWarpCfg::EX_BYTES;                                       // padded FFT exchange buffer
    static size_t warp_bytes(int N) {
        const size_t row = ((size_t)N * 8 + 127) / 128 * 128, own = STASH_BYTES + EX_BYTES;
        return row > own ? row : own;
    }
    // per block: the queries' cut-off words and bound arrays (4 KB), the pass twiddles (31 x 32 cf = 8 KB) and the split
    // twiddles (512 cf = 4 KB) -- every second stage reads both tables, and with ~200 KB of shared memory configured the L1
    // is too small to hold them
    static constexpr size_t TW_BYTES = 31 * 32 * sizeof(cf) + 64, SW_BYTES = (size_t)(G::M / 2) * sizeof(cf);
    static size_t table_bytes() { return (size_t)QMAX * 2 * sizeof(void *) + TW_BYTES + SW_BYTES; }
    static int warps(int N) {
        const size_t w = (SMEM_BUDGET - table_bytes()) / warp_bytes(N);
        return (int)(w > MAX_WARPS ? MAX_WARPS : w);
    }
    static size_t smem_bytes(int N) { return (size_t)warps(N) * warp_bytes(N) + table_bytes(); }
};

template <int NZ>
__global__ void __launch_bounds__(RefineMultiCfg::MAX_WARPS * 32, 1)
refine_multi_kernel(const ScreenParams prm, const MultiQuery *__restrict__ queries, const int nq, const unsigned warp_bytes,
                    unsigned *__restrict__ next_series) {
    using C = RefineMultiCfg;
    using G = typename C::G;
    constexpr int P = 32, M = G::M, QI = C::QI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[C::MAX_WARPS];

    const int w = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    const int nwarps = (int)(blockDim.x >> 5);
    const int count = (int)prm.count;
    // series are handed out by a global counter (the cost of a series varies with the number of its second stages by two
    // orders of magnitude: a fixed assignment left most warps idle behind the unlucky ones); the counter starts at the number
    // of warps of the grid, whose first series is their own index
    const int pos0 = (int)(blockIdx.x * nwarps) + w;
    unsigned char *buf = smem_raw + (size_t)w * warp_bytes;
    const cd *rowc = reinterpret_cast<const cd *>(buf);
    cf *stash = reinterpret_cast<cf *>(buf);                                  // slot j of lane t at [32 j + t]
    cf *ex = reinterpret_cast<cf *>(buf + C::STASH_BYTES);                    // this warp's FFT exchange buffer
    unsigned **s_cut = reinterpret_cast<unsigned **>(smem_raw + (size_t)nwarps * warp_bytes);      // [QMAX] cut-off words
    float **s_out = reinterpret_cast<float **>(s_cut + C::QMAX);                                   // [QMAX] bound arrays
    cf *s_tw = reinterpret_cast<cf *>(s_out + C::QMAX);                                            // pass twiddles W_1024^(j t)
    cf *s_sw = reinterpret_cast<cf *>(reinterpret_cast<unsigned char *>(s_tw) + C::TW_BYTES);      // split twiddles exp(-2 pi i k / n)
    const int N = prm.N;
    const int Nh = (N + 1) >> 1;      // complex slots holding samples (odd N: the pad column of the last one holds the row's mean, RowStat)
    const unsigned bar = smem_u32(&bars[w]);
    const int partner = (P - t) & (P - 1);
    const bool lane0 = (t == 0);
    const bool last_in = t + (NZ - 1) * 32 < Nh;
    const int nqi = (nq + 31) >> 5;

    if (t == 0) {
        mbar_init(bar, 1);
        if (pos0 < count) bulk_load(smem_u32(buf), prm.slab + (int64_t)pos0 * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
    }
    for (int q = threadIdx.x; q < C::QMAX; q += blockDim.x) {
        s_cut[q] = q < nq ? queries[q].cut : nullptr;
        s_out[q] = q < nq ? queries[q].out_U : nullptr;
    }
    for (int i = threadIdx.x; i < 31 * 32; i += blockDim.x) s_tw[i] = prm.twp[i];
    for (int i = threadIdx.x; i < M / 2; i += blockDim.x) {
        const float4 s = prm.sw[i];
        s_sw[i] = cf{s.x, s.y};
    }
    __syncthreads();

    unsigned phase = 0;
    int next = 0;
    for (int pos = pos0; pos < count; pos = next, phase ^= 1u) {
        // this series' bounds against all queries and the queries' running cut-offs: in flight under the transform
        float U[QI];
        unsigned cutr[QI];
#pragma unroll
        for (int i = 0; i < QI; i++) {
            U[i] = -1.f;
            cutr[i] = 0x7f800000u;          // +inf: no such query
            if (i < nqi && t + 32 * i < nq) {
                U[i] = s_out[t + 32 * i][pos];
                cutr[i] = ld_relaxed_u32(s_cut[t + 32 * i]);
            }
        }
        const RowStat rs = prm.row_stat[pos];
        const double mu = rs.mean;
        mbar_wait(bar, phase);

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
        __syncwarp();                       // the row is in registers: its buffer becomes exchange + stash
        if (t == 0) next = (int)atomicAdd(next_series, 1u);
        next = __shfl_sync(0xffffffffu, next, 0);

        Dft32Lead<NZ, float>::run(v);
#pragma unroll
        for (int j = 0; j < P; j++) {
            cf val = v[Perm<P>::at(j)];
            if (j > 0) val = cmul(val, s_tw[(j - 1) * 32 + t]);
            ex[G::pad(32 * t + j)] = val;
        }
        __syncwarp();
        fft_pass_load<10, 5, 1, float>(v, ex, t);
        Dft<P, float>::run(v);                          // v[Perm(j)] = Z[t + 32*j]
#pragma unroll
        for (int j = 0; j < P; j++) stash[32 * j + t] = v[Perm<P>::at(j)];      // every lane reads back only what it wrote

        // which (query group i, lane) pairs take the second stage: one ballot per group, then ONE rolled loop over the
        // pairs -- the second stage is ~45 KB of code, and a copy per group (the loop unrolled) thrashed the instruction
        // cache (measured: 10x slower)
        unsigned todo[QI];
#pragma unroll
        for (int i = 0; i < QI; i++) todo[i] = __ballot_sync(0xffffffffu, U[i] >= __uint_as_float(cutr[i]) && U[i] < 1.5f);
        unsigned changed = 0u;
#pragma unroll 1
        while (true) {
            // next pair: the lowest group with a bit left (warp-uniform select chains: the arrays stay in registers)
            int gi = -1;
            unsigned gm = 0u;
#pragma unroll
            for (int i = QI - 1; i >= 0; i--)
                if (todo[i]) {
                    gi = i;
                    gm = todo[i];
                }
            if (gi < 0) break;
            const int ql = __ffs(gm) - 1;
            float Usel = 0.f;
            unsigned csel = 0u;
#pragma unroll
            for (int i = 0; i < QI; i++) {
                if (i == gi) {
                    todo[i] = gm & (gm - 1);
                    Usel = U[i];
                    csel = cutr[i];
                }
            }
            {
                float Uq = __shfl_sync(0xffffffffu, Usel, ql);
                const float cut_now = __uint_as_float(__shfl_sync(0xffffffffu, csel, ql));
                const MultiQuery mq = queries[32 * gi + ql];
                cf z[P];
#pragma unroll
                for (int j = 0; j < P; j++) z[Perm<P>::at(j)] = stash[32 * j + t];
#pragma unroll
                for (int j = 0; j < P / 2; j++) {
                    const cf zk = z[Perm<P>::at(j)];
                    const cf zp = z[Perm<P>::at(P - 1 - j)];
                    const cf zs = z[Perm<P>::at((P - j) & (P - 1))];
                    cf src, zm;
                    src.x = lane0 ? zs.x : zp.x;
                    src.y = lane0 ? zs.y : zp.y;
                    zm.x = __shfl_sync(0xffffffffu, src.x, partner);
                    zm.y = __shfl_sync(0xffffffffu, src.y, partner);
                    const float4 x = mq.sx[t + 32 * j];
                    cf ok, om;
                    pointwise_pair(zk, zm, s_sw[t + 32 * j], cf{x.x, x.y}, cf{x.z, x.w}, ok, om);
                    cf rcv;
                    rcv.x = __shfl_sync(0xffffffffu, om.x, partner);
                    rcv.y = __shfl_sync(0xffffffffu, om.y, partner);
                    z[Perm<P>::at(j)] = ok;
                    cf &hi = z[Perm<P>::at(P - 1 - j)];
                    hi.x = lane0 ? hi.x : rcv.x;
                    hi.y = lane0 ? hi.y : rcv.y;
                    if (j > 0) {
                        cf &own = z[Perm<P>::at(P - j)];
                        own.x = lane0 ? om.x : own.x;
                        own.y = lane0 ? om.y : own.y;
                    }
                }
                {
                    cf &midz = z[Perm<P>::at(P / 2)];
                    cf ok, om;
                    pointwise_pair(midz, midz, cf{0.f, -1.f}, mq.x_mid, mq.x_mid, ok, om);
                    midz.x = lane0 ? ok.x : midz.x;
                    midz.y = lane0 ? ok.y : midz.y;
                }
                cf u[P];
#pragma unroll
                for (int j = 0; j < P; j++) u[j] = z[Perm<P>::at(j)];
                Dft<P, float>::run(u);
                __syncwarp();                           // the previous use of the exchange buffer has been read
#pragma unroll
                for (int j = 0; j < P; j++) {
                    cf val = u[Perm<P>::at(j)];
                    if (j > 0) val = cmul(val, s_tw[(j - 1) * 32 + t]);
                    ex[G::pad(32 * t + j)] = val;
                }
                __syncwarp();
                fft_pass_load<10, 5, 1, float>(u, ex, t);
                Dft<P, float>::run(u);
                float m_in = 0.f, m_out = 0.f;
                const int base = 2 * t - prm.win_lo;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const cf r = u[Perm<P>::at(j)];
                    const bool in0 = ((base + 64 * j) & (2 * M - 1)) <= prm.win_len;
                    const bool in1 = ((base + 64 * j + 1) & (2 * M - 1)) <= prm.win_len;
                    const float b0 = fabsf(r.y), b1 = fabsf(r.x);
                    m_in = fmaxf(m_in, in0 ? b0 : 0.f);
                    m_out = fmaxf(m_out, in0 ? 0.f : b0);
                    m_in = fmaxf(m_in, in1 ? b1 : 0.f);
                    m_out = fmaxf(m_out, in1 ? 0.f : b1);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, off));
                    m_out = fmaxf(m_out, __shfl_xor_sync(0xffffffffu, m_out, off));
                }
                float L = -1.f;
                Uq = refine_decide(Uq, m_in * rs.rstd, m_out * rs.rstd, L, 0);
                if (t == 0) atomicAdd(reinterpret_cast<unsigned long long *>(mq.cut + 2), 1ull);
                if (L >= prm.thr && L >= cut_now) cut_count_and_raise_at(mq.cut, prm.top_n, L, t);
#pragma unroll
                for (int i = 0; i < QI; i++)
                    if (i == gi && t == ql) {
                        U[i] = Uq;
                        changed |= 1u << i;
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < QI; i++)
            if (changed & (1u << i)) s_out[t + 32 * i][pos] = U[i];
        // the stash is no longer needed: hand the buffer to the copy engine for the warp's next row
        __syncwarp();
        if (t == 0 && next < count) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(smem_u32(buf), prm.slab + (int64_t)next * prm.ld, (unsigned)(N + (N & 1)) * 8u, bar);
        }
    }
}

#endif  // __CUDACC__

}  // namespace muse
