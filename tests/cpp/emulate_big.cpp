// CPU emulation of score_screen_big_kernel's per-thread phases (muse_screen_big.cuh compiled as host code).
// Each barrier-delimited phase is run for all T "threads" of a series in sequence.  Checked against a direct
// evaluation in long double:
//   * the bound  sum_f |Y_f| A_f  against  (1/n) sum_f |Y_f||X_f|  computed from an O(n^2)-free reference
//     (long-double DFT by recursion is not needed: |Y_f| comes from a plain double FFT written here);
//   * the second stage's maxima of |cc| inside / outside the lag window against the direct correlation
//     cc[k] = sum_t x'p[(t+k) mod n] * yp[t] (SURVEY section 8a closed form) of the CENTRED, un-normalised series.
// Prints one line per configuration; exit code 0 iff every configuration is within tolerance.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "muse_screen_big.cuh"

using namespace muse;
typedef std::complex<double> zd;
static const long double PI_L = 3.14159265358979323846264338327950288L;

static void fft_rec(std::vector<zd> &a, bool inv) {
    const size_t n = a.size();
    if (n == 1) return;
    std::vector<zd> e(n / 2), o(n / 2);
    for (size_t i = 0; i < n / 2; i++) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    fft_rec(e, inv);
    fft_rec(o, inv);
    for (size_t k = 0; k < n / 2; k++) {
        const long double ang = (inv ? 2 : -2) * PI_L * (long double)k / (long double)n;
        const zd w((double)cosl(ang), (double)sinl(ang));
        a[k] = e[k] + w * o[k];
        a[k + n / 2] = e[k] - w * o[k];
    }
}

template <int LOG2M>
static int run_case(int N, unsigned seed, int max_lag) {
    using C = ScreenBigCfg<LOG2M>;
    constexpr int M = C::M, n = 2 * M, T = C::T;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> ref(N), y(N);
    for (int i = 0; i < N; i++) {
        ref[i] = 0.1 * U(rng) + (i >= N / 2 - 5 && i < N / 2 + 5 ? 1.5 : 0.0);
        y[i] = 7.0 + 0.1 * U(rng) + (i >= N / 2 + 31 && i < N / 2 + 40 ? 2.5 : 0.0);
    }
    const int Nh = (N + 1) / 2;      // an odd N: the last slot holds (y[N-1], pad), and the pad is centred to exactly 0
    // ---- the reference side in double: x' = znorm(ref)/(N-1), zeros TRAIL (the kernel's rotation), Xt = X/(2n) ----
    double rm = 0, ym = 0;
    for (int i = 0; i < N; i++) { rm += ref[i]; ym += y[i]; }
    rm /= N; ym /= N;
    double rs = 0, ys = 0;
    for (int i = 0; i < N; i++) { rs += (ref[i] - rm) * (ref[i] - rm); ys += (y[i] - ym) * (y[i] - ym); }
    const double rsd = std::sqrt(rs / (N - 1)), ysd = std::sqrt(ys / (N - 1));
    std::vector<zd> X(n), Yc(n);
    for (int i = 0; i < n; i++) {
        X[i] = i < N ? (ref[i] - rm) / rsd / (N - 1) : 0.0;
        Yc[i] = i < N ? (y[i] - ym) : 0.0;
    }
    std::vector<zd> xt = X, yt = Yc;
    fft_rec(xt, false);
    fft_rec(yt, false);
    // direct answer: cc' = ifft(conj(Y) X) (un-normalised by std), bound = (1/n) sum |Y||X|
    std::vector<zd> cc(n);
    long double bound = 0;
    for (int f = 0; f < n; f++) { cc[f] = std::conj(yt[f]) * xt[f]; bound += (long double)std::abs(yt[f]) * std::abs(xt[f]); }
    bound /= n;
    fft_rec(cc, true);
    for (auto &c : cc) c /= (double)n;
    // window in the rotated index (zeros trail): true lag index = idx (no pad rotation in this harness: both trail)
    // kernel convention: (idx - win_lo) mod n <= win_len
    const int win_lo = ((-max_lag) % n + n) % n, win_len = 2 * max_lag;
    double w_in = 0, w_out = 0;
    for (int k = 0; k < n; k++) {
        const bool in = (((k - win_lo) % n + n) % n) <= win_len;
        const double a = std::fabs(cc[k].real());
        if (in) w_in = std::max(w_in, a); else w_out = std::max(w_out, a);
    }
    // ---- tables as the library builds them ----
    std::vector<cf> twi(C::TW_TOTAL);
    fill_big_twiddles(LOG2M, twi.data(), [](long long num, long long den) {
        return cf{(float)cosl(-2 * PI_L * num / den), (float)sinl(-2 * PI_L * num / den)};
    });
    const std::vector<cf> &twp = twi;
    std::vector<float4> sw(M / 2), sx(M / 2);
    auto Xt = [&](int k) { return xt[k] / (2.0 * n); };
    for (int k = 0; k < M / 2; k++) {
        const double wa = std::abs(Xt(k)) * (k == 0 ? 1.0 : 2.0), wc = std::abs(Xt(M - k)) * (k == 0 ? 1.0 : 2.0);
        sw[k] = make_float4((float)cosl(-2 * PI_L * k / n), (float)sinl(-2 * PI_L * k / n), (float)wa, (float)wc);
        sx[k] = make_float4((float)Xt(k).real(), (float)Xt(k).imag(), (float)Xt(M - k).real(), (float)Xt(M - k).imag());
    }
    const float a_mid = (float)(std::abs(Xt(M / 2)) * 2.0);
    const cf x_mid{(float)Xt(M / 2).real(), (float)Xt(M / 2).imag()};

    // ---- the kernel's phases, thread by thread ----
    std::vector<std::vector<cf>> regs(T, std::vector<cf>(32));
    std::vector<cf> sm(C::SM_ELEMS);
    for (int t = 0; t < T; t++)
        for (int j = 0; j < 32; j++) {
            const int e = t + j * T;
            regs[t][j] = e < Nh ? cf{(float)(y[2 * e] - ym), 2 * e + 1 < N ? (float)(y[2 * e + 1] - ym) : 0.f} : cf{0.f, 0.f};
        }
    for (int t = 0; t < T; t++) big_fwd_pass0<LOG2M>(regs[t].data(), sm.data(), t, twp.data());
    for (int t = 0; t < T; t++) big_load_stride_t<LOG2M>(regs[t].data(), sm.data(), t);
    for (int t = 0; t < T; t++) big_fwd_pass1<LOG2M>(regs[t].data(), sm.data(), t, twp.data());
    for (int t = 0; t < T; t++) big_fwd_last<LOG2M>(regs[t].data(), sm.data(), t);
    double acc = 0;
    for (int t = 0; t < T; t++) acc += big_split_bound<LOG2M>(regs[t].data(), t, sw.data(), a_mid);
    // the pair bound the kernel takes: sum over mirror pairs of sqrt(|2Y_k|^2 + |2Y_(M-k)|^2) sqrt(A_k^2 + A_(M-k)^2), + bin M/2
    std::vector<float> sb(M / 2);
    long double pair_want = 0;
    for (int k = 0; k < M / 2; k++) {
        const long double wa = sw[k].z, wc = sw[k].w;
        sb[k] = (float)std::sqrt((double)(2.0L * (wa * wa + wc * wc))) * (1.f + 1e-6f);
        const long double yk = 2.0L * std::abs(yt[k]), ym2 = 2.0L * std::abs(yt[M - k]);     // |2Y_k|, |2Y_(M-k)| (k = 0: DC, Nyquist)
        pair_want += std::sqrt((double)(yk * yk + ym2 * ym2)) * std::sqrt((double)(wa * wa + wc * wc));
    }
    pair_want += 2.0L * std::abs(yt[M / 2]) * a_mid;
    double acc_pair = 0;
    for (int t = 0; t < T; t++) acc_pair += big_split_bound<LOG2M, true>(regs[t].data(), t, sw.data(), a_mid, sb.data());
    for (int t = 0; t < T; t++) big_pointwise<LOG2M>(regs[t].data(), t, sw.data(), sx.data(), x_mid);
    for (int t = 0; t < T; t++) big_inv_pass0<LOG2M>(regs[t].data(), sm.data(), t, twi.data());
    for (int t = 0; t < T; t++) big_inv_pass1_load<LOG2M>(regs[t].data(), sm.data(), t);
    for (int t = 0; t < T; t++) big_inv_pass1<LOG2M>(regs[t].data(), sm.data(), t, twi.data());
    for (int t = 0; t < T; t++) big_load_stride_t<LOG2M>(regs[t].data(), sm.data(), t);
    float k_in = 0, k_out = 0;
    for (int t = 0; t < T; t++) {
        Dft<32, float>::run(regs[t].data());
        float a, b;
        big_window_max<LOG2M>(regs[t].data(), t, win_lo, win_len, a, b);
        k_in = std::max(k_in, a);
        k_out = std::max(k_out, b);
    }
    const double e_bound = std::fabs(acc - (double)bound) / ysd;
    const double e_in = std::fabs(k_in - w_in) / ysd, e_out = std::fabs(k_out - w_out) / ysd;
    const double e_pair = std::fabs(acc_pair - (double)pair_want) / ysd;
    const bool ok = e_bound < 2e-5 && e_in < 2e-5 && e_out < 2e-5 && e_pair < 2e-5 && acc_pair >= acc * (1.0 - 1e-6);
    printf("pair bound %.6f (err %.2e)  ", (double)pair_want / ysd, e_pair);
    printf("n=%5d N=%5d max_lag=%4d  bound %.6f (err %.2e)  in %.6f (err %.2e)  out %.6f (err %.2e)  %s\n", n, N, max_lag,
           (double)bound / ysd, e_bound, w_in / ysd, e_in, w_out / ysd, e_out, ok ? "ok" : "FAIL");
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run_case<11>(2050, 1, 30);
    bad += run_case<11>(4096, 2, 2047);
    bad += run_case<11>(3000, 3, 0);
    bad += run_case<12>(4100, 4, 100);
    bad += run_case<12>(8192, 5, 240);
    bad += run_case<13>(10080, 6, 240);
    bad += run_case<13>(16384, 7, 60);
    bad += run_case<13>(8194, 8, 5);
    bad += run_case<12>(4101, 9, 17);
    bad += run_case<13>(10081, 10, 240);
    return bad ? 1 : 0;
}
