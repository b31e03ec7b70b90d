// muse_fft.cuh -- register-resident radix-2..32 DFTs and the per-thread pieces of the
// shared-memory Stockham FFT used by every score kernel.
//
// Everything here is __host__ __device__ so that the exact same index arithmetic is
// exercised on the CPU (tests/emulate_kernel.cpp runs the per-thread phases for all
// "threads" of a series in sequence, standing in for the barriers) before any GPU
// time is spent.
//
// Replaces gonum dsp/fourier's Coefficients/Sequence as used by xcorr.go:183-186.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define MUSE_HD __host__ __device__ __forceinline__
#else
#define MUSE_HD inline
#endif

namespace muse {

template <typename F>
struct alignas(2 * sizeof(F)) cx {   // naturally aligned: one 64-/128-bit load or store per element
    F x, y;
};

template <typename F> MUSE_HD cx<F> cadd(cx<F> a, cx<F> b) { return cx<F>{a.x + b.x, a.y + b.y}; }
template <typename F> MUSE_HD cx<F> csub(cx<F> a, cx<F> b) { return cx<F>{a.x - b.x, a.y - b.y}; }
template <typename F> MUSE_HD cx<F> cmul(cx<F> a, cx<F> b) {
    return cx<F>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename F> MUSE_HD cx<F> cconj(cx<F> a) { return cx<F>{a.x, -a.y}; }
template <typename F> MUSE_HD cx<F> cmul_negi(cx<F> a) { return cx<F>{a.y, -a.x}; }   // a * (-i)
template <typename F> MUSE_HD cx<F> cmul_i(cx<F> a) { return cx<F>{-a.y, a.x}; }      // a * (+i)

// ---- fp32 complex arithmetic on packed pairs ------------------------------------------
// sm_100 has two-wide fp32 instructions (PTX add/sub/mul/fma .f32x2 -> SASS FADD2/FMUL2/
// FFMA2) whose operands take a free half swap (.LO_HI), per-half negation (.NP) and scalar
// broadcast (.F32): a complex add is ONE instruction, a complex multiply TWO, and
// conj / *(+-i) fold into the operand modifiers of the consumer.  The (re, im) pair of a
// cx<float> is the packed operand; ptxas keeps it in an aligned register pair.  Each half is
// an IEEE round-to-nearest fp32 operation, so the error analysis of the scalar code holds.
// Host builds (tests/cpp) use the scalar forms.
#if defined(__CUDA_ARCH__)
namespace f32x2 {
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ cx<float> up(u64 r) {
    cx<float> v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ cx<float> add(cx<float> a, cx<float> b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(d);
}
__device__ __forceinline__ cx<float> mul(cx<float> a, cx<float> b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(d);
}
__device__ __forceinline__ cx<float> fma(cx<float> a, cx<float> b, cx<float> c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)), "l"(pk(c.x, c.y)));
    return up(d);
}
}  // namespace f32x2
#endif

MUSE_HD cx<float> cadd(cx<float> a, cx<float> b) {
#if defined(__CUDA_ARCH__)
    return f32x2::add(a, b);
#else
    return cx<float>{a.x + b.x, a.y + b.y};
#endif
}
MUSE_HD cx<float> csub(cx<float> a, cx<float> b) {
#if defined(__CUDA_ARCH__)
    return f32x2::add(a, cx<float>{-b.x, -b.y});
#else
    return cx<float>{a.x - b.x, a.y - b.y};
#endif
}
// a*b = a * (b.x, b.x) + (-a.y, a.x) * (b.y, b.y): FMUL2 + FFMA2, b's halves broadcast
MUSE_HD cx<float> cmul(cx<float> a, cx<float> b) {
#if defined(__CUDA_ARCH__)
    const cx<float> t = f32x2::mul(a, cx<float>{b.x, b.x});
    return f32x2::fma(cx<float>{-a.y, a.x}, cx<float>{b.y, b.y}, t);
#else
    return cx<float>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
#endif
}
// component-wise a*b and a*b + c on the two halves (not complex products)
MUSE_HD cx<float> pmul(cx<float> a, cx<float> b) {
#if defined(__CUDA_ARCH__)
    return f32x2::mul(a, b);
#else
    return cx<float>{a.x * b.x, a.y * b.y};
#endif
}
MUSE_HD cx<float> pfma(cx<float> a, cx<float> b, cx<float> c) {
#if defined(__CUDA_ARCH__)
    return f32x2::fma(a, b, c);
#else
    return cx<float>{a.x * b.x + c.x, a.y * b.y + c.y};
#endif
}

// cos(2*pi*k/32), k = 0..8, correctly rounded
MUSE_HD constexpr double cos32_q(int k) {
    return k == 0 ? 1.0
         : k == 1 ? 0.98078528040323044912618223613423903697393373089333609500291
         : k == 2 ? 0.92387953251128675612818318939678828682241662586364248611509
         : k == 3 ? 0.83146961230254523707878837761790575673856081198797360312956
         : k == 4 ? 0.70710678118654752440084436210484903928483593768847403658834
         : k == 5 ? 0.55557023301960222474283081394853287437493719075480404592415
         : k == 6 ? 0.38268343236508977172845998403039886676134456248562704143380
         : k == 7 ? 0.19509032201612826784828486847702224092769161775195480775450
         : 0.0;
}
// cos(2*pi*k/32) for any k in [0, 32)
MUSE_HD constexpr double cos32(int k) {
    return k <= 8 ? cos32_q(k) : k <= 16 ? -cos32_q(16 - k) : k <= 24 ? -cos32_q(k - 16) : cos32_q(32 - k);
}
MUSE_HD constexpr double sin32(int k) { return cos32((k + 24) & 31); }   // sin(a) = cos(a - pi/2)

// v *= W_N^K = exp(-2*pi*i*K/N), N | 32, K compile-time.
template <int N, int K, typename F>
MUSE_HD void twiddle_const(cx<F> &v) {
    constexpr int k32 = ((K % N) * (32 / N)) & 31;
    if (k32 == 0) return;
    if (k32 == 8) { v = cmul_negi(v); return; }
    if (k32 == 16) { v = cx<F>{-v.x, -v.y}; return; }
    if (k32 == 24) { v = cmul_i(v); return; }
    constexpr F c = (F)cos32(k32), s = (F)(-sin32(k32));
    v = cmul(v, cx<F>{c, s});
}

// ---- in-register forward DFTs (exp(-i)); results land at v[base + Perm<N>(k)] ----
template <int N> struct Perm { MUSE_HD static constexpr int at(int k) { return k; } };
template <> struct Perm<8> { MUSE_HD static constexpr int at(int k) { return 4 * (k & 1) + (k >> 1); } };
template <> struct Perm<16> { MUSE_HD static constexpr int at(int k) { return 4 * (k & 3) + (k >> 2); } };
template <> struct Perm<32> { MUSE_HD static constexpr int at(int k) { return 8 * (k & 3) + Perm<8>::at(k >> 2); } };

template <typename F> MUSE_HD void dft2(cx<F> &a, cx<F> &b) {
    cx<F> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <typename F> MUSE_HD void dft4(cx<F> &a, cx<F> &b, cx<F> &c, cx<F> &d) {
    cx<F> t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = cmul_negi(csub(b, d));
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, t3);
    d = csub(t1, t3);
}

// dft4 whose last 4 - NNZ inputs are known to be zero (pruned first stage of a zero-padded
// transform): 6 complex additions for NNZ == 3, 4 for NNZ == 2
template <int NNZ, typename F> MUSE_HD void dft4_lead(cx<F> &a, cx<F> &b, cx<F> &c, cx<F> &d) {
    static_assert(NNZ >= 2 && NNZ <= 4, "dft4_lead: 2..4 leading non-zero inputs");
    if constexpr (NNZ == 4) {
        dft4(a, b, c, d);
    } else if constexpr (NNZ == 3) {
        const cx<F> t0 = cadd(a, c), t1 = csub(a, c), t3 = cmul_negi(b), bb = b;
        a = cadd(t0, bb);
        c = csub(t0, bb);
        b = cadd(t1, t3);
        d = csub(t1, t3);
    } else {
        const cx<F> aa = a, bb = b, t3 = cmul_negi(b);
        a = cadd(aa, bb);
        c = csub(aa, bb);
        b = cadd(aa, t3);
        d = csub(aa, t3);
    }
}

template <int N, typename F> struct Dft;
template <typename F> struct Dft<1, F> { MUSE_HD static void run(cx<F> *) {} };
template <typename F> struct Dft<2, F> { MUSE_HD static void run(cx<F> *v) { dft2(v[0], v[1]); } };
template <typename F> struct Dft<4, F> { MUSE_HD static void run(cx<F> *v) { dft4(v[0], v[1], v[2], v[3]); } };

// N = N1*N2: N2 inner DFTs of size N1 (stride N2), twiddle W_N^(n2*k1), N1 DFTs of size N2.
template <int N, int N1, int N2, typename F>
struct DftComposite {
    template <int n2, int k1> MUSE_HD static void tw(cx<F> *v) {
        twiddle_const<N, n2 * k1>(v[n2 + N2 * k1]);
        if constexpr (k1 + 1 < N1) tw<n2, k1 + 1>(v);
    }
    template <int n2> MUSE_HD static void inner(cx<F> *v) {
        if constexpr (N1 == 2) dft2(v[n2], v[n2 + N2]);
        else dft4(v[n2], v[n2 + N2], v[n2 + 2 * N2], v[n2 + 3 * N2]);
        if constexpr (n2 > 0) tw<n2, 1>(v);
        if constexpr (n2 + 1 < N2) inner<n2 + 1>(v);
    }
    template <int k1> MUSE_HD static void outer(cx<F> *v) {
        Dft<N2, F>::run(v + N2 * k1);
        if constexpr (k1 + 1 < N1) outer<k1 + 1>(v);
    }
    MUSE_HD static void run(cx<F> *v) {
        inner<0>(v);
        outer<0>(v);
    }
};
template <typename F> struct Dft<8, F> { MUSE_HD static void run(cx<F> *v) { DftComposite<8, 2, 4, F>::run(v); } };
template <typename F> struct Dft<16, F> { MUSE_HD static void run(cx<F> *v) { DftComposite<16, 4, 4, F>::run(v); } };
template <typename F> struct Dft<32, F> { MUSE_HD static void run(cx<F> *v) { DftComposite<32, 4, 8, F>::run(v); } };

// 32-point DFT whose inputs v[NZ..31] are zero (NZ >= 16): the radix-4 first stage is pruned.
template <int NZ, typename F>
struct Dft32Lead {
    static_assert(NZ >= 16 && NZ <= 32, "Dft32Lead: 16..32 leading non-zero inputs");
    using D = DftComposite<32, 4, 8, F>;
    template <int n2> MUSE_HD static void inner(cx<F> *v) {
        constexpr int left = NZ - n2;                        // k1 with n2 + 8*k1 < NZ
        constexpr int cnt = left >= 25 ? 4 : left >= 17 ? 3 : 2;
        dft4_lead<cnt>(v[n2], v[n2 + 8], v[n2 + 16], v[n2 + 24]);
        if constexpr (n2 > 0) D::template tw<n2, 1>(v);
        if constexpr (n2 + 1 < 8) inner<n2 + 1>(v);
    }
    MUSE_HD static void run(cx<F> *v) {
        inner<0>(v);
        D::template outer<0>(v);
    }
};

// ---- Stockham pass geometry ------------------------------------------------------
// An M = 2^LOG2M point complex FFT done by T = M/P cooperating threads, P = 2^LOG2P
// points per thread.  Pass i has radix R_i = P except the last, which takes what is
// left.  Before pass i the sub-transform length is NCUR = M >> (LOG2P*i) and the
// stride is S = 1 << (LOG2P*i).  Butterfly b (0 <= b < M/R) has p = b / S, q = b % S,
// reads x[q + S*(p + j*NCUR/R)] and writes y[q + S*(R*p + j)] * W_NCUR^(j*p).
template <int LOG2M, int LOG2P>
struct Geo {
    static constexpr int M = 1 << LOG2M;
    static constexpr int P = 1 << LOG2P;
    static constexpr int T = M / P;
    static constexpr int LOG2T = LOG2M - LOG2P;
    static constexpr int NPASS = LOG2M == 0 ? 1 : (LOG2M + LOG2P - 1) / LOG2P;
    // shared-memory index padding: one element per P makes the stride of the first pass'
    // radix-P scatter (thread p writes P*p + j) odd in elements, which is conflict-free for
    // both 8-byte (half-warp) and 16-byte (quarter-warp) accesses
    static constexpr int LOG2PAD = LOG2P > 0 ? LOG2P : 4;
    static constexpr int MP = M + (M >> LOG2PAD);
    MUSE_HD static constexpr int pad(int i) { return i + (i >> LOG2PAD); }
    MUSE_HD static constexpr int log2r(int pass) {
        return (LOG2M - LOG2P * pass) < LOG2P ? (LOG2M - LOG2P * pass) : LOG2P;
    }
    // Per-pass twiddle tables, concatenated.  Pass i (not the last) owns (R-1) rows of
    // NCUR/R entries: row j-1, column p holds W_NCUR^(j*p).  With p = b >> LS and b the
    // butterfly index, lanes of a warp read consecutive columns in pass 0 (coalesced) and
    // a handful of distinct columns in later passes (broadcast) -- the flat W_M^k table
    // read at stride j cost up to 60 L1 wavefronts per load.
    MUSE_HD static constexpr int tw_cols(int pass) { return 1 << (LOG2M - LOG2P * pass - log2r(pass)); }
    MUSE_HD static constexpr int tw_size(int pass) {
        return (LOG2P * pass + log2r(pass) == LOG2M) ? 0 : ((1 << log2r(pass)) - 1) * tw_cols(pass);
    }
    MUSE_HD static constexpr int tw_off(int pass) { return pass == 0 ? 0 : tw_off(pass - 1) + tw_size(pass - 1); }
    static constexpr int TW_TOTAL = tw_off(NPASS);
};

// Host-side fill of the per-pass twiddle tables (TW has .x/.y): W_NCUR^(j*p) = exp(-2*pi*i*j*p/NCUR).
template <typename TW, typename FN>
inline void fill_pass_twiddles(int log2m, int log2p, TW *out, FN unit_root /* (num, den) -> TW */) {
    int off = 0;
    for (int pass = 0;; pass++) {
        const int ls = log2p * pass;
        if (ls >= log2m && !(log2m == 0 && pass == 0)) break;
        int lr = log2m - ls < log2p ? log2m - ls : log2p;
        if (ls + lr == log2m) break;   // last pass: no twiddles
        const int R = 1 << lr, ncur = 1 << (log2m - ls), cols = ncur / R;
        for (int j = 1; j < R; j++)
            for (int p = 0; p < cols; p++) out[off + (j - 1) * cols + p] = unit_root((long long)j * p, (long long)ncur);
        off += (R - 1) * cols;
    }
}

// One thread's share of one FFT pass: DFTs in registers, twiddles, scatter to smem.
// v[c*R + j] holds input j of butterfly b = t + c*T.  tw = the per-pass tables of
// Geo::tw_off / fill_pass_twiddles.
template <int LOG2M, int LOG2P, int PASS, typename F, typename TW>
MUSE_HD void fft_pass_compute_store(cx<F> *v, cx<F> *sm, int t, const TW *tw) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr int LS = LOG2P * PASS;
    constexpr int LR = G::log2r(PASS);
    constexpr int R = 1 << LR;
    constexpr int NB = G::P / R;
    constexpr bool LAST = (LS + LR == LOG2M);
    constexpr int COLS = G::tw_cols(PASS);
    const TW *twp = tw + G::tw_off(PASS);
#pragma unroll
    for (int c = 0; c < NB; c++) {
        Dft<R, F>::run(v + c * R);
        const int b = t + c * G::T;
        const int p = b >> LS;
        const int q = b & ((1 << LS) - 1);
        // Padded index of output j = pad(q + ((R*p + j) << LS)).  When j << LS is a multiple of the padding
        // period the index is (per-butterfly base) + (compile-time offset); in pass 0 with R equal to the
        // period it is (R + 1)*p + j.  The generic form costs a shift and two adds per element (a quarter
        // of the exact kernel's instructions were such index arithmetic).
        constexpr bool LIN = LS >= G::LOG2PAD;
        constexpr bool LIN0 = (LS == 0 && LR == G::LOG2PAD);
        const int base = LIN ? G::pad(q + ((R * p) << LS)) : (LIN0 ? (R + 1) * p : 0);
#pragma unroll
        for (int j = 0; j < R; j++) {
            cx<F> val = v[c * R + Perm<R>::at(j)];
            if (!LAST && j > 0) {
                const TW w = twp[(j - 1) * COLS + p];     // W_NCUR^(j*p)
                val = cmul(val, cx<F>{(F)w.x, (F)w.y});
            }
            const int idx = LIN ? base + (j << LS) + ((j << LS) >> G::LOG2PAD) : (LIN0 ? base + j : G::pad(q + ((R * p + j) << LS)));
            sm[idx] = val;
        }
    }
}

// Gather the inputs of pass PASS from smem into v (same layout as above).
template <int LOG2M, int LOG2P, int PASS, typename F>
MUSE_HD void fft_pass_load(cx<F> *v, const cx<F> *sm, int t) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr int LS = LOG2P * PASS;
    constexpr int LR = G::log2r(PASS);
    constexpr int R = 1 << LR;
    constexpr int NB = G::P / R;
    constexpr int LNR = LOG2M - LS - LR;   // log2(NCUR / R)
#pragma unroll
    for (int c = 0; c < NB; c++) {
        const int b = t + c * G::T;
        const int p = b >> LS;
        const int q = b & ((1 << LS) - 1);
        // pad(q + ((p + (j << LNR)) << LS)) = pad(q + (p << LS)) + (compile-time offset) when j << (LNR + LS)
        // is a multiple of the padding period
        constexpr bool LIN = (LNR + LS) >= G::LOG2PAD;
        const int base = G::pad(q + (p << LS));
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int idx = LIN ? base + (j << (LNR + LS)) + ((j << (LNR + LS)) >> G::LOG2PAD) : G::pad(q + ((p + (j << LNR)) << LS));
            v[c * R + j] = sm[idx];
        }
    }
}

// Natural-order index of register slot (c, j) after the LAST pass (p == 0, q == b).
template <int LOG2M, int LOG2P>
MUSE_HD constexpr int last_pass_index(int t, int c, int j) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr int PASS = G::NPASS - 1;
    constexpr int LS = LOG2P * PASS;
    return (t + c * G::T) + (j << LS);
}

// ---- real-FFT split, conj-multiply by X, and re-pack for the inverse ---------------
// Z = FFT_M(z), z[j] = y[2j] + i*y[2j+1].  With m = (M-k) mod M, w = exp(-2*pi*i*k/n):
//   e = Z[k] + conj(Z[m]),  o = -i*(Z[k] - conj(Z[m]))
//   2*Y[k] = e + w*o,  2*conj(Y[M-k]) = e - w*o
//   c_k = conj(2Y[k]) * Xt[k],  c_m = (e - w*o) * Xt[M-k]          (Xt = X / (2n))
//   e' = c_k + conj(c_m),  d' = c_k - conj(c_m),  o' = d' * conj(w)
//   Z'[k] = e' + i*o',  Z'[m] = conj(e') + i*conj(o')
// IFFT_M(Z') = cc[2j] + i*cc[2j+1] (already divided by n: xcorr.go:187).  The inverse
// is run as swap(FFT(swap(.))), so the swapped values are what gets stored.
template <typename F, typename TX>
MUSE_HD void pointwise_pair(cx<F> zk, cx<F> zm, cx<F> w, TX xk, TX xm, cx<F> &outk_swapped, cx<F> &outm_swapped) {
    cx<F> zmc = cconj(zm);
    cx<F> e = cadd(zk, zmc);
    cx<F> o = cmul_negi(csub(zk, zmc));
    cx<F> wo = cmul(w, o);
    cx<F> yk = cadd(e, wo);          // 2*Y[k]
    cx<F> ym = csub(e, wo);          // 2*conj(Y[M-k])
    cx<F> ck = cmul(cconj(yk), cx<F>{(F)xk.x, (F)xk.y});
    cx<F> cm = cmul(ym, cx<F>{(F)xm.x, (F)xm.y});
    cx<F> cmc = cconj(cm);
    cx<F> e2 = cadd(ck, cmc);
    cx<F> d2 = csub(ck, cmc);
    cx<F> o2 = cmul(d2, cconj(w));
    cx<F> zk2 = cadd(e2, cmul_i(o2));
    cx<F> zm2 = cadd(cconj(e2), cmul_i(cconj(o2)));
    outk_swapped = cx<F>{zk2.y, zk2.x};
    outm_swapped = cx<F>{zm2.y, zm2.x};
}

// Spectrum of the reference: Xt[k] = Y[k] * scale, Xt[M-k] = Y[M-k] * scale.
template <typename F>
MUSE_HD void untangle_pair(cx<F> zk, cx<F> zm, cx<F> w, F scale, cx<F> &yk_out, cx<F> &ymk_out) {
    cx<F> zmc = cconj(zm);
    cx<F> e = cadd(zk, zmc);
    cx<F> o = cmul_negi(csub(zk, zmc));
    cx<F> wo = cmul(w, o);
    cx<F> yk = cadd(e, wo);
    cx<F> ym = cconj(csub(e, wo));
    yk_out = cx<F>{yk.x * scale, yk.y * scale};
    ymk_out = cx<F>{ym.x * scale, ym.y * scale};
}

}  // namespace muse
