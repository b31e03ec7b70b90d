/*
 * CPU oracle (C) for go-muse's Batch.Run hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm, used (a) as the fast checker
 * for the CUDA path at sizes where the numpy oracle would take minutes and
 * (b) as bench.py's cpu_baseline / --impl reference arm ("kind": "port": the
 * reference is Go and there is no Go toolchain in this image, so go-muse itself
 * cannot be timed).  Nothing in the product path links or calls this file.
 *
 * Citations are file:line in /root/reference.  The arithmetic go-muse delegates
 * to gonum v0.7.0 (go.mod:8; source not vendored) is restated from the published
 * algorithms: floats.Sum = sequential sum, stat.StdDev = corrected two-pass
 * unbiased variance, dsp/fourier = real FFT with n/2+1 coefficients and an
 * un-normalised inverse.  The FFT here is a radix-4/2 Stockham complex FFT of
 * n/2 points with the usual real-input split; it agrees with FFTPACK to ~1e-15.
 * Pinned by tests/test_oracle_c.py against the numpy oracle and the reference's
 * own KATs (tests/golden/reference_kats.json).
 *
 * Threading mirrors muse_batch.go:104-128: one worker per label-group, each with
 * its own FFT scratch (muse_batch.go:62-64), a pthread pool standing in for the
 * goroutines + semaphore.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct { double re, im; } cpx;

/* xcorr.go:19-24 */
int64_t muse_oracle_next_pow_of2(double val) {
    if (val <= 0) return 0;
    return (int64_t)pow(2.0, ceil(log(val) / log(2.0)));
}

/* ---- plan: twiddles for an n-point real FFT (n power of two, n >= 2) ---- */
typedef struct {
    int64_t n, m;      /* real length, complex length n/2 */
    cpx *w;            /* w[k] = exp(-2*pi*i*k/m), k < m   (complex FFT)  */
    cpx *wr;           /* wr[k] = exp(-2*pi*i*k/n), k <= m/2.. (real split) */
} plan_t;

static plan_t *plan_new(int64_t n) {
    plan_t *p = (plan_t *)calloc(1, sizeof(plan_t));
    p->n = n; p->m = n / 2;
    p->w = (cpx *)malloc(sizeof(cpx) * (size_t)(p->m > 0 ? p->m : 1));
    p->wr = (cpx *)malloc(sizeof(cpx) * (size_t)(p->m + 1));
    for (int64_t k = 0; k < p->m; k++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)p->m;
        p->w[k].re = (double)cosl(a); p->w[k].im = (double)sinl(a);
    }
    for (int64_t k = 0; k <= p->m; k++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        p->wr[k].re = (double)cosl(a); p->wr[k].im = (double)sinl(a);
    }
    return p;
}
static void plan_free(plan_t *p) { if (p) { free(p->w); free(p->wr); free(p); } }

/* Stockham autosort complex FFT of m points; sign=-1 forward, +1 inverse
 * (un-normalised).  Result ends in x.  y is scratch of m. */
static void cfft(const plan_t *p, cpx *x, cpx *y, int sign) {
    int64_t m = p->m, n = m, s = 1;
    cpx *a = x, *b = y;
    while (n >= 4) {
        int64_t n1 = n / 4;
        for (int64_t q = 0; q < n1; q++) {
            /* twiddle index: exp(sign*2*pi*i*q/n) = w[q * (m/n)] (conj for inverse) */
            cpx w1 = p->w[(q * (m / n)) % m], w2 = p->w[(2 * q * (m / n)) % m], w3 = p->w[(3 * q * (m / n)) % m];
            if (sign > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
            for (int64_t r = 0; r < s; r++) {
                cpx A = a[r + s * (q)], B = a[r + s * (q + n1)], C = a[r + s * (q + 2 * n1)], D = a[r + s * (q + 3 * n1)];
                cpx apc = {A.re + C.re, A.im + C.im}, amc = {A.re - C.re, A.im - C.im};
                cpx bpd = {B.re + D.re, B.im + D.im}, bmd = {B.re - D.re, B.im - D.im};
                /* forward: -i*(b-d); inverse: +i*(b-d) */
                cpx jb = (sign < 0) ? (cpx){bmd.im, -bmd.re} : (cpx){-bmd.im, bmd.re};
                cpx t0 = {apc.re + bpd.re, apc.im + bpd.im};
                cpx t1 = {amc.re + jb.re, amc.im + jb.im};
                cpx t2 = {apc.re - bpd.re, apc.im - bpd.im};
                cpx t3 = {amc.re - jb.re, amc.im - jb.im};
                b[r + s * (4 * q + 0)] = t0;
                b[r + s * (4 * q + 1)] = (cpx){t1.re * w1.re - t1.im * w1.im, t1.re * w1.im + t1.im * w1.re};
                b[r + s * (4 * q + 2)] = (cpx){t2.re * w2.re - t2.im * w2.im, t2.re * w2.im + t2.im * w2.re};
                b[r + s * (4 * q + 3)] = (cpx){t3.re * w3.re - t3.im * w3.im, t3.re * w3.im + t3.im * w3.re};
            }
        }
        n = n1; s *= 4;
        cpx *t = a; a = b; b = t;
    }
    if (n == 2) {
        for (int64_t r = 0; r < s; r++) {
            cpx A = a[r], B = a[r + s];
            b[r] = (cpx){A.re + B.re, A.im + B.im};
            b[r + s] = (cpx){A.re - B.re, A.im - B.im};
        }
        cpx *t = a; a = b; b = t;
    }
    if (a != x) memcpy(x, a, sizeof(cpx) * (size_t)m);
}

/* fourier.FFT.Coefficients: seq[n] real -> coef[n/2+1].  z,scr: scratch of n/2. */
static void rfft_fwd(const plan_t *p, const double *seq, cpx *coef, cpx *z, cpx *scr) {
    int64_t m = p->m;
    if (p->n == 1) { coef[0] = (cpx){seq[0], 0}; return; }
    for (int64_t j = 0; j < m; j++) { z[j].re = seq[2 * j]; z[j].im = seq[2 * j + 1]; }
    cfft(p, z, scr, -1);
    for (int64_t k = 0; k <= m; k++) {
        cpx zk = z[k % m], zc = z[(m - k) % m];
        cpx e = {0.5 * (zk.re + zc.re), 0.5 * (zk.im - zc.im)};
        cpx o = {0.5 * (zk.im + zc.im), -0.5 * (zk.re - zc.re)};  /* (zk - conj(zc)) / (2i) */
        cpx w = p->wr[k];
        coef[k].re = e.re + (w.re * o.re - w.im * o.im);
        coef[k].im = e.im + (w.re * o.im + w.im * o.re);
    }
}

/* fourier.FFT.Sequence: coef[n/2+1] -> seq[n], UN-normalised. */
static void rfft_inv(const plan_t *p, const cpx *coef, double *seq, cpx *z, cpx *scr) {
    int64_t m = p->m;
    if (p->n == 1) { seq[0] = coef[0].re; return; }
    for (int64_t k = 0; k < m; k++) {
        cpx a = coef[k], b = coef[m - k];
        cpx e = {0.5 * (a.re + b.re), 0.5 * (a.im - b.im)};
        cpx d = {0.5 * (a.re - b.re), 0.5 * (a.im + b.im)};       /* (a - conj(b)) / 2 */
        cpx w = {p->wr[k].re, -p->wr[k].im};                       /* exp(+2*pi*i*k/n) */
        cpx o = {d.re * w.re - d.im * w.im, d.re * w.im + d.im * w.re};
        z[k].re = e.re - o.im;                                     /* e + i*o */
        z[k].im = e.im + o.re;
    }
    cfft(p, z, scr, +1);
    for (int64_t j = 0; j < m; j++) { seq[2 * j] = 2.0 * z[j].re; seq[2 * j + 1] = 2.0 * z[j].im; }
}

/* xcorr.go:84-95 with gonum's floats.Sum / stat.StdDev restated.  Returns 0 when
 * std == 0 (errStdDevZero), else 1; z receives the normalised series. */
static int z_normalize(const double *x, int64_t N, double *z) {
    double sum = 0;
    for (int64_t i = 0; i < N; i++) sum += x[i];
    double mu = sum / (double)N;
    for (int64_t i = 0; i < N; i++) z[i] = x[i] - mu;
    double s2 = 0;
    for (int64_t i = 0; i < N; i++) s2 += z[i];
    double mean = s2 / (double)N, ss = 0, comp = 0;
    for (int64_t i = 0; i < N; i++) { double d = z[i] - mean; ss += d * d; comp += d; }
    double sd = sqrt((ss - comp * comp / (double)N) / (double)(N - 1));
    if (sd == 0) return 0;
    double inv = 1.0 / sd;
    for (int64_t i = 0; i < N; i++) z[i] *= inv;
    return 1;
}

typedef struct {
    plan_t *plan;
    int64_t N, n;
    cpx *X;            /* muse_batch.go:47: rfft(zeroPad(znorm(ref)/(N-1), n)) */
} batch_t;

typedef struct { double *seq, *zn; cpx *coef, *z, *scr; } scratch_t;

static scratch_t scratch_new(int64_t N, int64_t n) {
    scratch_t s;
    s.seq = (double *)malloc(sizeof(double) * (size_t)n);
    s.zn = (double *)malloc(sizeof(double) * (size_t)N);
    s.coef = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    s.z = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    s.scr = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    return s;
}
static void scratch_free(scratch_t *s) { free(s->seq); free(s->zn); free(s->coef); free(s->z); free(s->scr); }

/* muse_batch.go:35-47.  Returns NULL when std(ref)==0 or n is not a power of two. */
static batch_t *batch_new(const double *ref, int64_t N) {
    int64_t n = muse_oracle_next_pow_of2((double)N);
    if (n < 1 || (n & (n - 1))) return NULL;
    batch_t *b = (batch_t *)calloc(1, sizeof(batch_t));
    b->N = N; b->n = n; b->plan = plan_new(n);
    b->X = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    scratch_t s = scratch_new(N, n);
    int ok = z_normalize(ref, N, s.zn);
    if (ok) {
        double sc = 1.0 / (double)(N - 1);
        for (int64_t i = 0; i < n - N; i++) s.seq[i] = 0;
        for (int64_t i = 0; i < N; i++) s.seq[n - N + i] = s.zn[i] * sc;
        rfft_fwd(b->plan, s.seq, b->X, s.z, s.scr);
    }
    scratch_free(&s);
    if (!ok) { plan_free(b->plan); free(b->X); free(b); return NULL; }
    return b;
}
static void batch_free(batch_t *b) { if (b) { plan_free(b->plan); free(b->X); free(b); } }

/* xcorr.go:160-197: returns lag, *mv = signed peak; cc_out (n doubles) optional. */
static int64_t xcorr_with_x(const batch_t *b, const double *y, scratch_t *s, double *mv, double *cc_out) {
    int64_t N = b->N, n = b->n, m = n / 2;
    if (!z_normalize(y, N, s->zn)) { *mv = 0; if (cc_out) memset(cc_out, 0, sizeof(double) * (size_t)n); return 0; }
    for (int64_t i = 0; i < n - N; i++) s->seq[i] = 0;                 /* :176-178 */
    for (int64_t i = 0; i < N; i++) s->seq[n - N + i] = s->zn[i];     /* :179-181 */
    rfft_fwd(b->plan, s->seq, s->coef, s->z, s->scr);                 /* :183 */
    for (int64_t k = 0; k <= m; k++) {                                /* :184-185 conj, mult */
        double cr = s->coef[k].re, ci = -s->coef[k].im;
        s->coef[k].re = cr * b->X[k].re - ci * b->X[k].im;
        s->coef[k].im = cr * b->X[k].im + ci * b->X[k].re;
    }
    rfft_inv(b->plan, s->coef, s->seq, s->z, s->scr);                 /* :186 */
    double inv = 1.0 / (double)n;
    int64_t mi = 0; double best = 0;                                  /* :39-50 */
    for (int64_t i = 0; i < n; i++) {
        s->seq[i] *= inv;                                             /* :187 */
        if (fabs(s->seq[i]) > fabs(best)) { best = s->seq[i]; mi = i; }
    }
    if (cc_out) memcpy(cc_out, s->seq, sizeof(double) * (size_t)n);
    *mv = s->seq[mi];
    if (mi > n / 2) mi -= n;                                          /* :192-194 */
    return mi;
}

/* ------------------------------ exported API ------------------------------ */

/* Full cc vector of one pair (KAT support).  Returns 0 ok, 1 std(ref)==0, 2 std(y)==0. */
int muse_oracle_xcorr_with_x(const double *ref, const double *y, int64_t N, double *cc, int64_t *lag, double *mv) {
    batch_t *b = batch_new(ref, N);
    if (!b) return 1;
    scratch_t s = scratch_new(N, b->n);
    int zero = !z_normalize(y, N, s.zn);
    *lag = xcorr_with_x(b, y, &s, mv, cc);
    scratch_free(&s); batch_free(b);
    return zero ? 2 : 0;
}

/* ---- minimal pthread parallel-for (dynamic chunks off an atomic counter) ---- */
typedef void (*range_fn)(void *ctx, int64_t lo, int64_t hi, scratch_t *s);
typedef struct { range_fn fn; void *ctx; int64_t total, chunk; atomic_llong next; int64_t N, n; } pf_t;

static void *pf_worker(void *arg) {
    pf_t *p = (pf_t *)arg;
    scratch_t s = scratch_new(p->N, p->n);                             /* muse_batch.go:62-64 */
    for (;;) {
        int64_t lo = (int64_t)atomic_fetch_add(&p->next, (long long)p->chunk);
        if (lo >= p->total) break;
        int64_t hi = lo + p->chunk < p->total ? lo + p->chunk : p->total;
        p->fn(p->ctx, lo, hi, &s);
    }
    scratch_free(&s);
    return NULL;
}

int muse_oracle_max_threads(void) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

static void parallel_for(int64_t total, int64_t chunk, int nthreads, int64_t N, int64_t n, range_fn fn, void *ctx) {
    if (nthreads <= 0) nthreads = muse_oracle_max_threads();
    if (nthreads > 1024) nthreads = 1024;
    pf_t p; p.fn = fn; p.ctx = ctx; p.total = total; p.chunk = chunk; p.N = N; p.n = n;
    atomic_init(&p.next, 0);
    if (nthreads == 1 || total <= chunk) { pf_worker(&p); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++) if (pthread_create(&th[started], NULL, pf_worker, &p) == 0) started++;
    pf_worker(&p);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
}

typedef struct { const batch_t *b; const double *Y; int signed_scores; double *scores; int64_t *lags; } score_ctx;

static void score_range(void *vc, int64_t lo, int64_t hi, scratch_t *s) {
    score_ctx *c = (score_ctx *)vc;
    for (int64_t i = lo; i < hi; i++) {
        double mv; int64_t lag = xcorr_with_x(c->b, c->Y + i * c->b->N, s, &mv, NULL);
        if (c->signed_scores) { if (mv > 1.0) mv = 1.0; else if (mv < -1.0) mv = -1.0; }  /* muse.go:72-76 */
        else { mv = fabs(mv); if (mv > 1.0) mv = 1.0; }                                   /* muse_batch.go:74-77 */
        c->scores[i] = mv; c->lags[i] = lag;
    }
}

/* Per-series unsigned clamped score (muse_batch.go:74-77) or signed clamped
 * score (muse.go:72-76) and lag for every row of Y[S][N]. */
int muse_oracle_score_all(const double *ref, int64_t N, const double *Y, int64_t S, int signed_scores,
                          double *scores, int64_t *lags, int nthreads) {
    batch_t *b = batch_new(ref, N);
    if (!b) return 1;
    score_ctx c = {b, Y, signed_scores, scores, lags};
    parallel_for(S, 64, nthreads, N, b->n, score_range, &c);
    batch_free(b);
    return 0;
}

typedef struct { double score; int64_t lag, idx; } rep_t;

static int rep_cmp(const void *a, const void *b) {
    const rep_t *x = (const rep_t *)a, *y = (const rep_t *)b;
    if (fabs(x->score) > fabs(y->score)) return -1;
    if (fabs(x->score) < fabs(y->score)) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

typedef struct { const batch_t *b; const double *Y; const int64_t *members, *goff; rep_t *reps; } group_ctx;

/* muse_batch.go:56-93 scoreSingle for groups [lo, hi) */
static void group_range(void *vc, int64_t lo, int64_t hi, scratch_t *s) {
    group_ctx *c = (group_ctx *)vc;
    for (int64_t g = lo; g < hi; g++) {
        rep_t best = {0, 0, -1};
        for (int64_t j = c->goff[g]; j < c->goff[g + 1]; j++) {
            int64_t i = c->members[j];
            double mv; int64_t lag = xcorr_with_x(c->b, c->Y + i * c->b->N, s, &mv, NULL);
            mv = fabs(mv); if (mv > 1.0) mv = 1.0;                     /* :74-77 */
            if (mv > best.score || best.idx < 0) { best.score = mv; best.lag = lag; best.idx = i; }  /* :87 */
        }
        c->reps[g] = best;
    }
}

/* Batch.Run + Results.Fetch for one Run on a fresh Results (muse_batch.go:99-130,
 * results.go:46-87).  members[] lists series indices grouped by group:
 * group g owns members[goff[g] .. goff[g+1]).  One worker per group
 * (muse_batch.go:116-122).  Ties: first member wins in a group (:87), lowest
 * series index wins at the top-N boundary (one of the orders Go's map iteration
 * can produce).  Returns 0 ok, 1 invalid reference. */
int muse_oracle_batch_run(const double *ref, int64_t N, const double *Y, int64_t S,
                          const int64_t *members, const int64_t *goff, int64_t G,
                          int64_t max_lag, int64_t top_n, double threshold, int sign_filter,
                          double *out_scores, int64_t *out_lags, int64_t *out_idx, int64_t *n_out,
                          int nthreads) {
    (void)S;
    batch_t *b = batch_new(ref, N);
    if (!b) return 1;
    rep_t *reps = (rep_t *)malloc(sizeof(rep_t) * (size_t)(G > 0 ? G : 1));
    group_ctx gc = {b, Y, members, goff, reps};
    parallel_for(G, 16, nthreads, N, b->n, group_range, &gc);
    int64_t k = 0;
    for (int64_t g = 0; g < G; g++) {                                  /* results.go:46-59 */
        rep_t r = reps[g];
        if (r.idx < 0) continue;
        int ok = (double)llabs(r.lag) <= (double)max_lag && fabs(r.score) >= threshold &&
                 (sign_filter == 0 || (r.score > 0 && sign_filter == 1) || (r.score < 0 && sign_filter == -1));
        if (ok) reps[k++] = r;
    }
    qsort(reps, (size_t)k, sizeof(rep_t), rep_cmp);                    /* results.go:62-66,81-85 */
    if (k > top_n) k = top_n > 0 ? top_n : 0;
    for (int64_t i = 0; i < k; i++) { out_scores[i] = reps[i].score; out_lags[i] = reps[i].lag; out_idx[i] = reps[i].idx; }
    *n_out = k;
    free(reps); batch_free(b);
    return 0;
}
