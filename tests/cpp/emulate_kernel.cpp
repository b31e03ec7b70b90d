// CPU emulation of the fused score kernel's per-thread phases (muse_score.cuh /
// muse_fft.cuh compiled as host code).  Each barrier-delimited phase is run for all
// T "threads" of a series in sequence; the result is compared with a direct
// O(n^2) long-double evaluation of cc[k] = sum_t x'p[(t+k) mod n] * yp[t]
// (SURVEY section 8a closed form).  Prints one line per configuration; exit code 0
// iff every configuration is within tolerance.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "muse_score.cuh"

using namespace muse;

template <int LOG2M, int LOG2P, int PASS, typename F>
struct Passes {
    static void run(std::vector<std::vector<cx<F>>> &regs, std::vector<cx<F>> &sm, const std::vector<cd> &twM,
                    bool first_from_regs) {
        using G = Geo<LOG2M, LOG2P>;
        if (!(PASS == 0 && first_from_regs))
            for (int t = 0; t < G::T; t++) fft_pass_load<LOG2M, LOG2P, PASS, F>(regs[t].data(), sm.data(), t);
        constexpr bool LAST = (PASS == G::NPASS - 1);
        if (!LAST) {
            for (int t = 0; t < G::T; t++)
                fft_pass_compute_store<LOG2M, LOG2P, PASS, F, cd>(regs[t].data(), sm.data(), t, twM.data());
            Passes<LOG2M, LOG2P, LAST ? PASS : PASS + 1, F>::run(regs, sm, twM, false);
        }
    }
};

// Runs the FFT; on return the LAST pass inputs are loaded into regs but not yet transformed.
template <int LOG2M, int LOG2P, typename F>
static void fft_all_but_last(std::vector<std::vector<cx<F>>> &regs, std::vector<cx<F>> &sm,
                             const std::vector<cd> &twM, bool first_from_regs) {
    Passes<LOG2M, LOG2P, 0, F>::run(regs, sm, twM, first_from_regs);
}

template <int LOG2M, int LOG2P, typename F>
static void last_pass_to_smem(std::vector<std::vector<cx<F>>> &regs, std::vector<cx<F>> &sm, const std::vector<cd> &twM) {
    using G = Geo<LOG2M, LOG2P>;
    for (int t = 0; t < G::T; t++)
        fft_pass_compute_store<LOG2M, LOG2P, G::NPASS - 1, F, cd>(regs[t].data(), sm.data(), t, twM.data());
}

template <int LOG2M, int LOG2P, typename F>
static void last_pass_in_regs(std::vector<std::vector<cx<F>>> &regs) {
    using G = Geo<LOG2M, LOG2P>;
    constexpr int PASS = G::NPASS - 1;
    constexpr int R = 1 << G::log2r(PASS);
    for (int t = 0; t < G::T; t++)
        for (int c = 0; c < G::P / R; c++) Dft<R, F>::run(regs[t].data() + c * R);
}

static const long double PI_L = 3.14159265358979323846264338327950288L;

template <int LOG2M, int LOG2P, typename F>
static int run_case(int N, unsigned seed, double tol, bool even_path) {
    using G = Geo<LOG2M, LOG2P>;
    const int M = G::M, n = 2 * M, T = G::T, P = G::P;
    if (N > n || N <= n / 2 && n > 2) { /* N must satisfy nextPow2(N) == n */ }
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> ref(N), y(N + 2);
    for (int i = 0; i < N; i++) { ref[i] = U(rng) + (i > N / 3 && i < N / 3 + 5 ? 3.0 : 0.0); y[i] = 50.0 + U(rng) + (i > N / 2 && i < N / 2 + 4 ? 2.5 : 0.0); }
    // tables
    std::vector<cd> twM(G::TW_TOTAL + 1), twn(M / 2 + 1);
    fill_pass_twiddles(LOG2M, LOG2P, twM.data(), [](long long num, long long den) {
        return cd{(double)cosl(-2 * PI_L * num / den), (double)sinl(-2 * PI_L * num / den)};
    });
    for (int k = 0; k <= M / 2; k++) twn[k] = cd{(double)cosl(-2 * PI_L * k / n), (double)sinl(-2 * PI_L * k / n)};

    // ---- reference spectrum through the same phases (MODE_REF of the kernel) ----
    std::vector<std::vector<cd>> regs(T, std::vector<cd>(P));
    std::vector<cd> sm(G::MP + 1);
    std::vector<cd> Xt(M + 1);
    double ref_sd;
    {
        double sum = 0;
        for (int t = 0; t < T; t++) { double s; if (even_path) load_row<LOG2M, LOG2P, true>(regs[t].data(), ref.data(), N, t, s); else load_row<LOG2M, LOG2P, false>(regs[t].data(), ref.data(), N, t, s); sum += s; }
        double mu = sum / N, ss = 0, comp = 0;
        for (int t = 0; t < T; t++) { double a, c; center_row<LOG2M, LOG2P>(regs[t].data(), N, t, mu, a, c); ss += a; comp += c; }
        ref_sd = std::sqrt((ss - comp * comp / N) / (N - 1));
        fft_all_but_last<LOG2M, LOG2P, double>(regs, sm, twM, true);
        last_pass_to_smem<LOG2M, LOG2P, double>(regs, sm, twM);
        const double scale = 1.0 / (4.0 * n * ref_sd * (N - 1));
        for (int k = 0; k <= M / 2; k++) {
            int m = (M - k) & (M - 1);
            cd a, b;
            untangle_pair(sm[G::pad(k)], sm[G::pad(m)], twn[k], scale, a, b);
            Xt[M - k] = b;
            Xt[k] = a;
        }
    }
    // ---- series ----
    double sum = 0;
    for (int t = 0; t < T; t++) { double s; if (even_path) load_row<LOG2M, LOG2P, true>(regs[t].data(), y.data(), N, t, s); else load_row<LOG2M, LOG2P, false>(regs[t].data(), y.data(), N, t, s); sum += s; }
    double mu = sum / N, ss = 0, comp = 0;
    for (int t = 0; t < T; t++) { double a, c; center_row<LOG2M, LOG2P>(regs[t].data(), N, t, mu, a, c); ss += a; comp += c; }
    std::vector<std::vector<cx<F>>> fr(T, std::vector<cx<F>>(P));
    std::vector<cx<F>> fsm(G::MP + 1);
    for (int t = 0; t < T; t++) for (int r = 0; r < P; r++) fr[t][r] = cx<F>{(F)regs[t][r].x, (F)regs[t][r].y};
    fft_all_but_last<LOG2M, LOG2P, F>(fr, fsm, twM, true);
    last_pass_to_smem<LOG2M, LOG2P, F>(fr, fsm, twM);
    for (int t = 0; t < T; t++) pointwise_phase<LOG2M, LOG2P, F, cd>(fsm.data(), t, Xt.data(), twn.data());
    fft_all_but_last<LOG2M, LOG2P, F>(fr, fsm, twM, false);
    last_pass_in_regs<LOG2M, LOG2P, F>(fr);
    Peak pk{0.0, 0.0, 0x7fffffff};
    std::vector<double> cc(n, 0.0);
    for (int t = 0; t < T; t++) {
        Peak q = argmax_local<LOG2M, LOG2P, F>(fr[t].data(), t);
        peak_merge(pk, q.a, q.v, q.idx);
        constexpr int PASS = G::NPASS - 1;
        constexpr int R = 1 << G::log2r(PASS);
        for (int c = 0; c < P / R; c++)
            for (int j = 0; j < R; j++) {
                int e = last_pass_index<LOG2M, LOG2P>(t, c, j);
                cx<F> val = fr[t][c * R + Perm<R>::at(j)];
                cc[2 * e] = val.y;
                cc[2 * e + 1] = val.x;
            }
    }
    double score; int lag;
    finish_series(pk, ss, comp, N, n, false, score, lag);
    const double sd = std::sqrt((ss - comp * comp / N) / (N - 1));

    // ---- direct long-double evaluation ----
    std::vector<long double> xp(n, 0.0L), yp(n, 0.0L);
    long double mr = 0, my = 0;
    for (int i = 0; i < N; i++) { mr += ref[i]; my += y[i]; }
    mr /= N; my /= N;
    long double vr = 0, vy = 0;
    for (int i = 0; i < N; i++) { vr += (ref[i] - mr) * (ref[i] - mr); vy += (y[i] - my) * (y[i] - my); }
    long double sr = sqrtl(vr / (N - 1)), sy = sqrtl(vy / (N - 1));
    for (int i = 0; i < N; i++) { xp[n - N + i] = (ref[i] - mr) / sr / (N - 1); yp[n - N + i] = (y[i] - my) / sy; }
    double maxerr = 0, best = 0; int bi = 0;
    for (int k = 0; k < n; k++) {
        long double acc = 0;
        for (int t = 0; t < n; t++) acc += xp[(t + k) % n] * yp[t];
        double got = cc[k] / sd;
        maxerr = std::fmax(maxerr, std::fabs(got - (double)acc));
        if (std::fabs((double)acc) > std::fabs(best) + 1e-12) { best = (double)acc; bi = k; }
    }
    int want_lag = bi > n / 2 ? bi - n : bi;
    double want_score = std::fmin(std::fabs(best), 1.0);
    bool ok = maxerr <= tol && std::fabs(score - want_score) <= tol && lag == want_lag;
    printf("LOG2M=%d LOG2P=%d N=%d n=%d %s %s: max|cc err|=%.3e score=%.12f want=%.12f lag=%d want=%d %s\n", LOG2M, LOG2P, N,
           n, sizeof(F) == 8 ? "f64" : "f32", even_path ? "vec" : "scalar", maxerr, score, want_score, lag, want_lag, ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    // (LOG2M, LOG2P) instantiations the library uses: P = min(16, M) for fp64
    bad += run_case<0, 0, double>(2, 1, 1e-12, true);
    bad += run_case<1, 1, double>(4, 2, 1e-12, true);
    bad += run_case<1, 1, double>(3, 3, 1e-12, false);
    bad += run_case<2, 2, double>(8, 4, 1e-12, true);
    bad += run_case<2, 2, double>(5, 5, 1e-12, false);
    bad += run_case<3, 3, double>(12, 6, 1e-12, true);
    bad += run_case<4, 4, double>(31, 7, 1e-12, false);
    bad += run_case<5, 4, double>(50, 8, 1e-12, true);
    bad += run_case<6, 4, double>(100, 9, 1e-12, true);
    bad += run_case<7, 4, double>(255, 10, 1e-12, false);
    bad += run_case<8, 4, double>(480, 11, 1e-12, true);
    bad += run_case<9, 4, double>(1000, 12, 1e-12, true);
    bad += run_case<10, 4, double>(1440, 13, 1e-12, true);
    bad += run_case<10, 5, double>(1440, 14, 1e-12, true);
    bad += run_case<10, 5, float>(1440, 15, 2e-5, true);
    bad += run_case<8, 4, float>(480, 16, 2e-5, true);
    bad += run_case<11, 4, double>(2500, 17, 1e-12, true);
    printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
    return bad ? 1 : 0;
}
