// kernels_long.cu -- series whose FFT length does not fit one thread block's shared memory (n > 16384), and
// the FFT form of the generic xCorr.
//
// go-muse has no length limit: NewBatch takes any series (muse_batch.go:23-52, n = nextPowOf2(length)) and
// BenchmarkXCorrWithX runs 16385 samples at n = 32768 (xcorr_test.go:330-348).  Above n = 16384 the fused
// kernels of muse_exact.cuh stop (8192 fp64 complex points are the shared memory of one SM), so these rows go
// through a Stockham FFT whose passes stream through global memory, a chunk of series at a time; the working
// set of a chunk (2 x 16 n bytes per PAIR of series) is sized to stay in the 126 MB L2 for moderate n.
//
// Two real series share one complex transform.  With z = y1 + i*y2 (both z-normalised rows with their LEADING
// zero pad, xcorr.go:176-181) and Z = FFT_n(z), the correlation theorem for a real reference x gives
//     cc1 + i*cc2 = IFFT_n( X[k] * Z[(n-k) mod n] ),        X = FFT_n(x)
// because conj(Y1[k]) + i*conj(Y2[k]) = Y1[n-k] + i*Y2[n-k] = Z[n-k] for real y1, y2: no real/imaginary split,
// one forward and one inverse complex transform per pair, the real part of the result is series 1's
// cross-correlation (xcorr.go:183-187) and the imaginary part series 2's.  The inverse runs as
// swap(FFT(swap(.))).  Rows stay unscaled by 1/std until the end (finish_series, muse_score.cuh), as in the
// fused kernel.
#include <algorithm>
#include <cstring>

#include "muse_launch.h"

namespace muse {

// tw[k] = exp(-2*pi*i*k/n), k < n.  sincospi of the exactly representable -2k/n: within an ulp or two of the
// correctly rounded tables the short kernels get from the host's long double.
__global__ void __launch_bounds__(256) long_twiddle_kernel(cd *__restrict__ tw, long long n) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(-2.0 * (double)k / (double)n, &s, &c);
    tw[k] = cd{c, s};
}

__device__ __forceinline__ double long_block_sum(double v, double *red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) s += red[w];
    return s;
}

// Scratch of a chunk behind its two buffers (doubles): stats[4 per pair] = (ss, comp) of each series, then per pair and
// series LONG_NB_MAX partial row sums, 2 x LONG_NB_MAX partial (ss, comp), 3 x LONG_NB_MAX partial peaks.
constexpr int LONG_NB_MAX = 64;
constexpr int LONG_SCRATCH_PER_PAIR = 4 + 2 * 6 * LONG_NB_MAX;
__device__ __forceinline__ double *long_part(const LongParams &prm, long long pairs_cap, long long pair, int h, int what) {
    // what: 0 = row sums [NB], 1 = (ss, comp) [2 NB], 2 = peaks [3 NB]
    double *base = prm.stats + 4 * pairs_cap + (size_t)(2 * pair + h) * (6 * LONG_NB_MAX);
    return base + (what == 0 ? 0 : what == 1 ? LONG_NB_MAX : 3 * LONG_NB_MAX);
}
__device__ __forceinline__ const double *long_row(const LongParams &prm, long long first, long long pair, int h) {
    const long long pos = first + 2 * pair + h;
    if (pos >= prm.count) return nullptr;
    const long long row = prm.idx ? (long long)prm.idx[pos] : pos;
    return prm.slab + (size_t)row * (size_t)prm.ld;
}

// A row is cut into nb segments (blockIdx.x), one block each, so that a chunk of few long rows still fills the GPU; every
// sum is then taken over the segments' partial sums in segment order, so results do not depend on scheduling.
// Step 1: partial row sums (xcorr.go:85-86).
__global__ void __launch_bounds__(256)
long_sum_kernel(const LongParams prm, long long first, long long pairs_cap, int nb) {
    __shared__ double red[8];
    const long long pair = blockIdx.y;
    const long long N = prm.N, seg = (N + nb - 1) / nb;
    const long long lo = seg * blockIdx.x, hi = lo + seg < N ? lo + seg : N;
    for (int h = 0; h < 2; h++) {
        const double *row = long_row(prm, first, pair, h);
        if (!row) continue;            // uniform over the block
        double s = 0.0;
        for (long long i = lo + threadIdx.x; i < hi; i += 256) s += row[i];
        s = long_block_sum(s, red);
        if (threadIdx.x == 0) long_part(prm, pairs_cap, pair, h, 0)[blockIdx.x] = s;
    }
}

// Step 2: centred values with their partial sum of squares and the residual sum the corrected two-pass variance needs
// (xcorr.go:88), written as z[j] = (y1[j], y2[j]) behind n - N zeros.
__global__ void __launch_bounds__(256)
long_load_kernel(const LongParams prm, long long first, long long pairs_cap, int nb, cd *__restrict__ z) {
    __shared__ double red[8];
    const long long pair = blockIdx.y;
    const long long n = 1ll << prm.log2n;
    const long long N = prm.N, pad = n - N;
    cd *out = z + (size_t)pair * (size_t)n;
    const double *rows[2];
    double mu[2] = {0.0, 0.0};
    for (int h = 0; h < 2; h++) {
        rows[h] = long_row(prm, first, pair, h);
        if (!rows[h]) continue;
        const double *ps = long_part(prm, pairs_cap, pair, h, 0);
        double s = 0.0;
        for (int k = 0; k < nb; k++) s += ps[k];
        mu[h] = s / (double)N;
    }
    {   // this block's share of the leading zeros
        const long long zseg = (pad + nb - 1) / nb, zlo = zseg * blockIdx.x, zhi = zlo + zseg < pad ? zlo + zseg : pad;
        for (long long i = zlo + threadIdx.x; i < zhi; i += 256) out[i] = cd{0.0, 0.0};
    }
    const long long seg = (N + nb - 1) / nb;
    const long long lo = seg * blockIdx.x, hi = lo + seg < N ? lo + seg : N;
    double ss[2] = {0.0, 0.0}, cp[2] = {0.0, 0.0};
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        cd v{0.0, 0.0};
        if (rows[0]) {
            v.x = rows[0][i] - mu[0];
            ss[0] += v.x * v.x;
            cp[0] += v.x;
        }
        if (rows[1]) {
            v.y = rows[1][i] - mu[1];
            ss[1] += v.y * v.y;
            cp[1] += v.y;
        }
        out[pad + i] = v;
    }
    for (int h = 0; h < 2; h++) {
        if (!rows[h]) continue;
        const double a = long_block_sum(ss[h], red);
        const double c = long_block_sum(cp[h], red);
        if (threadIdx.x == 0) {
            double *pp = long_part(prm, pairs_cap, pair, h, 1);
            pp[2 * blockIdx.x] = a;
            pp[2 * blockIdx.x + 1] = c;
        }
    }
}

// Step 3: stats[2 * (2 pair + h)] = (ss, comp) from the partials, in segment order.
__global__ void __launch_bounds__(256)
long_stats_kernel(const LongParams prm, long long pairs, long long pairs_cap, int nb) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= 2 * pairs) return;
    const double *pp = long_part(prm, pairs_cap, i >> 1, (int)(i & 1), 1);
    double a = 0.0, c = 0.0;
    for (int k = 0; k < nb; k++) {
        a += pp[2 * k];
        c += pp[2 * k + 1];
    }
    prm.stats[2 * i] = a;
    prm.stats[2 * i + 1] = c;
}

// One Stockham pass of radix R over every transform of the chunk: butterfly j of a transform reads
// in[j + r n/R], multiplies by exp(-2*pi*i r (j mod Ns) / (Ns R)), takes the R-point DFT and writes
// out[(j / Ns) Ns R + (j mod Ns) + r Ns].  Ns = product of the radices of the passes before.
template <int R>
__global__ void __launch_bounds__(256)
long_pass_kernel(const cd *__restrict__ in, cd *__restrict__ out, const cd *__restrict__ tw, int log2n, int log2ns,
                 long long butterflies) {
    constexpr int LR = R == 8 ? 3 : (R == 4 ? 2 : 1);
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    if (g >= butterflies) return;
    const int lb = log2n - LR;                       // log2(butterflies per transform)
    const long long t = g >> lb, j = g & ((1ll << lb) - 1);
    const cd *src = in + ((size_t)t << log2n);
    cd *dst = out + ((size_t)t << log2n);
    const long long k = j & ((1ll << log2ns) - 1);
    const int shift = log2n - log2ns - LR;           // table step: n / (Ns R)
    cd v[R];
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = src[j + ((long long)r << lb)];
    if (log2ns > 0) {
        // W^(r k): one table read, the higher powers by multiplication (depth <= 3, a few ulp): the strided reads of a
        // flat table cost this pass twice the sectors of its data
        cd wp[R];
        wp[1] = tw[k << shift];
        if constexpr (R > 2) {
            wp[2] = cmul(wp[1], wp[1]);
            wp[3] = cmul(wp[2], wp[1]);
        }
        if constexpr (R > 4) {
            wp[4] = cmul(wp[2], wp[2]);
            wp[5] = cmul(wp[4], wp[1]);
            wp[6] = cmul(wp[3], wp[3]);
            wp[7] = cmul(wp[4], wp[3]);
        }
#pragma unroll
        for (int r = 1; r < R; r++) v[r] = cmul(v[r], wp[r]);
    }
    Dft<R, double>::run(v);
    const long long base = ((j >> log2ns) << (log2ns + LR)) + k;
#pragma unroll
    for (int r = 0; r < R; r++) dst[base + ((long long)r << log2ns)] = v[Perm<R>::at(r)];
}

// P[k] = swap(X[k] * Z[(n-k) mod n]); X carries the 1/n of xcorr.go:187.
__global__ void __launch_bounds__(256)
long_pointwise_kernel(const cd *__restrict__ Z, cd *__restrict__ P, const cd *__restrict__ X, int log2n, long long total) {
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    if (g >= total) return;
    const long long n = 1ll << log2n, t = g >> log2n, k = g & (n - 1);
    const cd z = Z[((size_t)t << log2n) + ((n - k) & (n - 1))];
    const cd p = cmul(X[k], z);
    P[g] = cd{p.y, p.x};
}

// Reference side (muse_batch.go:38-47): the transform of the centred reference row times 1/((N-1) std n).
__global__ void __launch_bounds__(256)
long_ref_kernel(const LongParams prm, const cd *__restrict__ Z) {
    const long long n = 1ll << prm.log2n;
    const double N = (double)prm.N;
    const double var = (prm.stats[0] - prm.stats[1] * prm.stats[1] / N) / (N - 1.0);
    const double sd = sqrt(var);
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k == 0) *prm.out_flag = (sd == 0.0) ? 1 : 0;
    if (k >= n) return;
    const double scale = 1.0 / (sd * (N - 1.0) * (double)n);
    prm.out_X[k] = cd{Z[k].x * scale, Z[k].y * scale};
}

// MODE_CC: the cross-correlation of the first series of the chunk, scaled by 1/std (xcorr.go:189-191 on the whole row)
__global__ void __launch_bounds__(256)
long_cc_kernel(const LongParams prm, const cd *__restrict__ W) {
    const long long n = 1ll << prm.log2n;
    const double N = (double)prm.N;
    const double var = (prm.stats[0] - prm.stats[1] * prm.stats[1] / N) / (N - 1.0);
    const double sd = sqrt(var);
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k == 0) *prm.out_flag = (sd == 0.0) ? 1 : 0;
    if (k >= n) return;
    prm.out_score[k] = sd == 0.0 ? 0.0 : W[k].y * (1.0 / sd);      // W = swap(cc1 + i cc2)
}

// maxAbsIndex over all n lags (xcorr.go:39-50, :189) for both series of a pair: partial peaks per segment of the lags ...
__global__ void __launch_bounds__(256)
long_peak_kernel(const LongParams prm, long long pairs_cap, int nb, const cd *__restrict__ W) {
    __shared__ double sa[2][256], sv[2][256];
    __shared__ int si[2][256];
    const long long pair = blockIdx.y;
    const long long n = 1ll << prm.log2n;
    const long long seg = (n + nb - 1) / nb, lo = seg * blockIdx.x, hi = lo + seg < n ? lo + seg : n;
    const cd *w = W + (size_t)pair * (size_t)n;
    Peak p0{0.0, 0.0, 0x7fffffff}, p1{0.0, 0.0, 0x7fffffff};
    for (long long j = lo + threadIdx.x; j < hi; j += 256) {
        const cd v = w[j];
        peak_merge(p0, fabs(v.y), v.y, (int)j);      // series 1: the real part of the un-swapped result
        peak_merge(p1, fabs(v.x), v.x, (int)j);
    }
    sa[0][threadIdx.x] = p0.a; sv[0][threadIdx.x] = p0.v; si[0][threadIdx.x] = p0.idx;
    sa[1][threadIdx.x] = p1.a; sv[1][threadIdx.x] = p1.v; si[1][threadIdx.x] = p1.idx;
    __syncthreads();
    if (threadIdx.x < 2) {
        const int h = threadIdx.x;
        Peak r{0.0, 0.0, 0x7fffffff};
        for (int i = 0; i < 256; i++) peak_merge(r, sa[h][i], sv[h][i], si[h][i]);
        double *pk = long_part(prm, pairs_cap, pair, h, 2) + 3 * blockIdx.x;
        pk[0] = r.a;
        pk[1] = r.v;
        pk[2] = (double)r.idx;
    }
}

// ... merged (lowest index wins ties, whatever the order), then finish_series.
__global__ void __launch_bounds__(256)
long_finish_kernel(const LongParams prm, long long first, long long pairs, long long pairs_cap, int nb) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= 2 * pairs) return;
    const long long pos = first + i;
    if (pos >= prm.count) return;
    const double *pk = long_part(prm, pairs_cap, i >> 1, (int)(i & 1), 2);
    Peak r{0.0, 0.0, 0x7fffffff};
    for (int k = 0; k < nb; k++) peak_merge(r, pk[3 * k], pk[3 * k + 1], (int)pk[3 * k + 2]);
    const long long row = prm.idx ? (long long)prm.idx[pos] : pos;
    double score;
    int lag;
    finish_series(r, prm.stats[2 * i], prm.stats[2 * i + 1], prm.N, 1 << prm.log2n, prm.signed_scores != 0, score, lag);
    prm.out_score[row] = score;
    prm.out_lag[row] = lag;
}

// every transform of `transforms` consecutive n-point rows of buf[0] -> returns the buffer index holding the result
static int long_fft(cd *buf[2], int cur, const cd *tw, int log2n, long long transforms, cudaStream_t st) {
    int log2ns = 0;
    auto pass = [&](int lr) {
        const long long bf = transforms << (log2n - lr);
        const unsigned blocks = (unsigned)((bf + 255) / 256);
        if (lr == 3) long_pass_kernel<8><<<blocks, 256, 0, st>>>(buf[cur], buf[cur ^ 1], tw, log2n, log2ns, bf);
        else if (lr == 2) long_pass_kernel<4><<<blocks, 256, 0, st>>>(buf[cur], buf[cur ^ 1], tw, log2n, log2ns, bf);
        else long_pass_kernel<2><<<blocks, 256, 0, st>>>(buf[cur], buf[cur ^ 1], tw, log2n, log2ns, bf);
        cur ^= 1;
        log2ns += lr;
    };
    if (log2n % 3) pass(log2n % 3);         // the short pass first: Ns = 1, no twiddles
    while (log2ns < log2n) pass(3);
    return cur;
}

size_t long_work_bytes(int log2n, long long chunk_pairs) {
    return 2 * sizeof(cd) * ((size_t)chunk_pairs << log2n) + sizeof(double) * LONG_SCRATCH_PER_PAIR * (size_t)chunk_pairs;
}

cudaError_t launch_long_twiddles(cd *tw, int log2n, cudaStream_t st) {
    const long long n = 1ll << log2n;
    long_twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tw, n);
    return cudaGetLastError();
}

// mode: MODE_SCORE / MODE_REF / MODE_CC of muse_exact.cuh.  work: long_work_bytes(log2n, chunk_pairs) bytes.
cudaError_t launch_long(int mode, LongParams p, void *work, long long chunk_pairs, cudaStream_t st) {
    if (p.log2n < 1 || p.log2n > 30 || chunk_pairs < 1) return cudaErrorInvalidValue;
    const long long n = 1ll << p.log2n;
    cd *buf[2];
    buf[0] = reinterpret_cast<cd *>(work);
    buf[1] = buf[0] + ((size_t)chunk_pairs << p.log2n);
    p.stats = reinterpret_cast<double *>(buf[1] + ((size_t)chunk_pairs << p.log2n));
    if (mode != MODE_SCORE) p.count = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (long long first = 0; first < p.count; first += 2 * chunk_pairs) {
        const long long pairs = std::min<long long>(chunk_pairs, (p.count - first + 1) / 2);
        // segments per row: enough blocks for two per SM, at least 4096 samples each
        int nb = (int)std::min<long long>(LONG_NB_MAX, std::max<long long>(1, (2 * sms + pairs - 1) / pairs));
        nb = (int)std::max<long long>(1, std::min<long long>(nb, (long long)p.N / 4096));
        const dim3 grid((unsigned)nb, (unsigned)pairs);
        const unsigned small = (unsigned)((2 * pairs + 255) / 256);
        long_sum_kernel<<<grid, 256, 0, st>>>(p, first, chunk_pairs, nb);
        long_load_kernel<<<grid, 256, 0, st>>>(p, first, chunk_pairs, nb, buf[0]);
        long_stats_kernel<<<small, 256, 0, st>>>(p, pairs, chunk_pairs, nb);
        int cur = long_fft(buf, 0, p.tw, p.log2n, pairs, st);
        if (mode == MODE_REF) {
            long_ref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, buf[cur]);
            return cudaGetLastError();
        }
        const long long total = pairs << p.log2n;
        long_pointwise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(buf[cur], buf[cur ^ 1], p.X, p.log2n, total);
        cur = long_fft(buf, cur ^ 1, p.tw, p.log2n, pairs, st);
        if (mode == MODE_CC) {
            long_cc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, buf[cur]);
            return cudaGetLastError();
        }
        long_peak_kernel<<<grid, 256, 0, st>>>(p, chunk_pairs, nb, buf[cur]);
        long_finish_kernel<<<small, 256, 0, st>>>(p, first, pairs, chunk_pairs, nb);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// ---- generic xCorr (xcorr.go:102-153) through the same passes ---------------------------------------
// xp, yp: the nn-long padded (and, when asked, z-normalised) rows of xcorr_prepare_kernel.  For a power-of-two nn the
// circular correlation is one L = nn transform; for any other nn it is the LINEAR correlation of the two rows
// (L >= 2 nn, trailing zeros) folded: cc[k] = lin[k] + lin[k - nn] = r[k] + r[L - nn + k].
__global__ void __launch_bounds__(256)
long_xy_pack_kernel(const double *__restrict__ xp, const double *__restrict__ yp, long long nn, long long L, cd *__restrict__ z) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= L) return;
    z[i] = i < nn ? cd{xp[i], yp[i]} : cd{0.0, 0.0};
}

// Z = FFT(xp + i yp): X[k] = (Z[k] + conj(Z[L-k]))/2, Y[k] = (Z[k] - conj(Z[L-k]))/(2i); P[k] = swap(X[k] conj(Y[k]))
__global__ void __launch_bounds__(256)
long_xy_pointwise_kernel(const cd *__restrict__ Z, cd *__restrict__ P, long long L) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= L) return;
    const cd a = Z[k], b = cconj(Z[(L - k) & (L - 1)]);
    const cd X = cd{0.5 * (a.x + b.x), 0.5 * (a.y + b.y)};
    const cd d = cd{0.5 * (a.x - b.x), 0.5 * (a.y - b.y)};      // i Y
    const cd Y = cmul_negi(d);
    const cd p = cmul(X, cconj(Y));
    P[k] = cd{p.y, p.x};
}

__global__ void __launch_bounds__(256)
long_xy_fold_kernel(const cd *__restrict__ W, long long nn, long long L, double scale, double *__restrict__ cc) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= nn) return;
    double r = W[k].y;                               // W = swap(r + i 0)
    if (L != nn) r += W[L - nn + k].y;
    cc[k] = r * scale / (double)L;
}

// work: 2 L complex values + L twiddles.  scale: what the direct kernel applies (1/(nn-1) for normalised inputs, 1 otherwise).
size_t long_xcorr_work_bytes(long long L) { return 3 * sizeof(cd) * (size_t)L; }

cudaError_t launch_long_xcorr(const double *xp, const double *yp, long long nn, long long L, double scale, double *cc, void *work,
                              cudaStream_t st) {
    int log2n = 0;
    while ((1ll << log2n) < L) log2n++;
    if ((1ll << log2n) != L || L < nn || (L != nn && L < 2 * nn)) return cudaErrorInvalidValue;
    cd *buf[2];
    buf[0] = reinterpret_cast<cd *>(work);
    buf[1] = buf[0] + L;
    cd *tw = buf[1] + L;
    const unsigned blocks = (unsigned)((L + 255) / 256);
    long_twiddle_kernel<<<blocks, 256, 0, st>>>(tw, L);
    long_xy_pack_kernel<<<blocks, 256, 0, st>>>(xp, yp, nn, L, buf[0]);
    int cur = long_fft(buf, 0, tw, log2n, 1, st);
    long_xy_pointwise_kernel<<<blocks, 256, 0, st>>>(buf[cur], buf[cur ^ 1], L);
    cur = long_fft(buf, cur ^ 1, tw, log2n, 1, st);
    long_xy_fold_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(buf[cur], nn, L, scale, cc);
    return cudaGetLastError();
}

}  // namespace muse
