// muse_api.cu -- the C ABI of include/muse_b200.h: handles, device memory, launches.
//
// Host code here only moves bytes, sizes launches and finishes the <= top_n sized tail
// (sorting the selected records); all arithmetic on series data happens in the kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/muse_b200.h"
#include "muse_launch.h"
#include "muse_select.cuh"
#include "muse_synth.cuh"
#include "muse_xcorr.cuh"

using namespace muse;

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError(); /* reported here: must not resurface in a later cudaGetLastError() check */ \
            return fail(e_ == cudaErrorMemoryAllocation ? MUSE_ERR_OUT_OF_MEMORY : MUSE_ERR_CUDA,  \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

extern "C" const char *muse_last_error(void) { return g_err; }
extern "C" const char *muse_version(void) { return "muse_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------
// Everything a batch needs per run whose size follows the store (not the reference): kept in a small
// per-context pool across muse_batch_create / muse_batch_destroy, because cudaMalloc / cudaFree /
// cudaHostAlloc of these cost milliseconds to seconds (measured: a NewBatch + Run + destroy cycle per
// request spent 3 s in cudaFree/cudaFreeHost once in four) while the run itself takes 2.7 ms.
#define MUSE_MAX_LABEL_KEYS 64   // label-key columns of a store (the facade adds one all-absent column)
#define MUSE_SCRATCH_POOL 272   // scratch sets kept per context: a multi-query launch has up to MUSE_MULTI_GROUP batches alive at once
#define MUSE_MULTI_GROUP 256    // queries of one launch of the tensor-core multi-query path (== TcCfg::TN == RefineMultiCfg::QMAX)
struct RunScratch {
    int64_t scratch_cap;
    double *d_score;
    int32_t *d_lag;
    int64_t *d_slot;
    unsigned long long *d_ckey, *d_skey;
    int32_t *d_cidx, *d_sidx;
    int32_t *d_clag, *d_slag;         // 2*lag + sign of the candidates (muse_select.cuh Cand)
    float *d_U;
    int32_t *d_list;
    float *d_L;                       // lower bounds (grouped screened runs, diagnostic entry point)
    signed char *d_W;                 // grouped screened runs: lag-window verdict of every refined series (ScreenParams::out_W)
    int64_t d_L_cap;
    unsigned char *h_pin;             // pinned mailbox for the small device->host results
    size_t h_pin_bytes;
    unsigned long long *d_counters;   // [0] ncand, [1] nselected, [2] exact list length, [3] refined
    SelectState *d_sel;
    int32_t *d_flag;
    unsigned *d_cut;                  // fused refinement state: [0] cut bits, [2..3] n_refined (u64), [4..] coarse + fine histogram
    // group table
    int64_t table_cap;
    unsigned long long *d_gmax, *d_hkeys;
    int32_t *d_gidx;
    // tables of the reference side.  tab_n: the FFT length the twiddle tables (twM, twn, twp_f, swtw) were filled
    // for -- a set that comes back from the pool with the same n keeps them, only the per-reference ones
    // (d_ref, Xt, sw_f, sx_f, d_mid) are rewritten by muse_batch_create
    int64_t tab_n, d_ref_cap;
    int tab_screen;                   // twp_f / swtw / sw_f / sx_f exist for tab_n
    double *d_ref;                    // padded copy of the reference row
    cd *Xt, *twM, *twn;
    cf *twp_f;                        // fp32 screening pass: per-pass twiddles
    cf *twi_f;                        // n = 4096 .. 16384: twiddle tables of muse_screen_big.cuh
    cf *twide_f;                      // n = 16384: twiddle tables of muse_screen_wide.cuh
    float2 *swtw;                     // split twiddles exp(-2*pi*i*k/n), k < M/2, fp32
    float4 *sw_f;                     // fused kernels: (twn, A[k], A[M-k]) per k < M/2
    float4 *sx_f;                     // fused refinement: (Xt[k], Xt[M-k]) in fp32 per k < M/2
    float *sb_f;                      // warp / sub-warp kernels: sqrt(2 (A[k]^2 + A[M-k]^2)) rounded up per k < M/2
    float *d_mid;                     // [0] std-zero flag of the reference (as float bits), [1] A[M/2], [2..3] Xt[M/2] in fp32
    cudaEvent_t ev[4];
    cudaStream_t aux;                 // the batch's own stream: the tails of the queries of a multi-query launch run side by side
};

struct muse_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;       // the stream in use
    cudaStream_t own_stream;   // created with the context
    std::mutex mu;
    std::vector<RunScratch> pool;   // scratch sets of destroyed batches, reused by the next muse_batch_create
    unsigned char *h_stage[3];      // pinned staging ring for rows that arrive in pageable host memory (muse_group_append)
    cudaEvent_t stage_ev[3];
    // multi-query bounds on the tensor cores (muse_bounds_tc.cuh): magnitudes of the store and weights of a launch's queries
    unsigned char *tc_a, *tc_b;     // bf16 tiles: [S/128][16][16 KB] and [16][32 KB]
    size_t tc_a_bytes;
    float *tc_mid, *tc_amid;        // [S] |2Y_(M/2)| per series, [256] A_q[M/2] per query
    int64_t tc_mid_cap;
    void **tc_ptrs;                 // device: [256] sw tables, [256] bound arrays
    unsigned *d_multi_next;         // work counter of refine_multi_kernel
    unsigned char *d_mt, *h_mt;     // parameter tables and results of a multi-query launch (MultiTables), device and pinned host
    int64_t mt_top_n;               // result records per query the two buffers were sized for
    cudaEvent_t mt_ev[5];           // stage boundaries of the last multi-query launch group: start, magnitudes, bounds, second stages, end
    int64_t multi_refined, multi_rescored;      // totals of the last muse_multi_run (second stages, fp64 re-scorings)
    void *d_multi_q;                // query table of score_screen_multi_kernel (ScreenMultiCfg::QC entries)
    double *d_multi_refs;           // [QC][d_multi_ld] reference rows of a multi-query launch, pad columns kept zero
    int64_t d_multi_ld, d_multi_n;
};

struct muse_group {
    muse_ctx *ctx;
    int64_t N;          // series length
    int64_t ld;         // row stride in doubles (multiple of 16 -> 128-byte rows)
    int64_t cap, size;
    double *slab;       // [cap][ld]
    int nkeys;
    int32_t *labels;    // [nkeys][cap]
    int32_t max_id[MUSE_MAX_LABEL_KEYS];
    int64_t global_offset;
    RowStat *row_stat;          // [cap] screening: fp64 mean and fp32 1/std of each row (filled lazily up to stats_upto)
    int64_t stats_cap, stats_upto;
};

struct muse_batch : RunScratch {
    muse_ctx *ctx;
    muse_group *g;
    int64_t N, n;
    int log2m;
    // fp32 screening pass (n = 128 .. 16384): the tables live in RunScratch
    int screen_ok;
    float a_mid;
    cf x_mid;
    muse_timing timing;
    int timing_pending;    // the last run was queued without a final synchronisation (muse_batch_run_partial_device)
    int fused_run;         // 1: score_fused, 2: score_fused_grouped (d_counters[2] = exact list length, [3] = refined)
    int prescreened;       // d_U and the cut-off state were filled by score_screen_multi_kernel: the next fused run starts at its tail
    int use_aux;           // queue this batch's work on its own stream (RunScratch::aux) instead of the context's
    int signed_run;        // the current run keeps the sign of the scores (muse.go:72-76)
    // n > MUSE_MAX_FUSED_FFT_LEN: Stockham passes through global memory (kernels_long.cu); Xt holds n entries, twM the n twiddles
    int is_long;
    void *d_long;          // work buffer of launch_long
    size_t d_long_bytes;
    long long long_pairs;  // pairs of series per chunk d_long was sized for
};

static inline cudaStream_t bstream(const muse_batch *b) { return b->use_aux ? b->aux : b->ctx->stream; }
// timing events of a run: not recorded for the batches of a multi-query launch (nobody reads their timings, and five
// calls per query are a fifth of what the host issues for one)
#define TIMING_EVENT(b, i, st)                             \
    do {                                                   \
        if (!(b)->use_aux) CU(cudaEventRecord((b)->ev[i], st)); \
    } while (0)

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {}
};

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
extern "C" int muse_ctx_create(int device, muse_ctx **out) {
    if (!out) return fail(MUSE_ERR_INVALID_ARG, "muse_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MUSE_ERR_NO_DEVICE, "no CUDA device (%s); muse_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(MUSE_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, count);
    CU(cudaSetDevice(device));
    muse_ctx *c = new muse_ctx();
    c->device = device;
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    *out = c;
    return MUSE_OK;
}

static void scratch_free(RunScratch &r) {
    cudaFree(r.d_score); cudaFree(r.d_lag); cudaFree(r.d_slot);
    cudaFree(r.d_ckey); cudaFree(r.d_skey); cudaFree(r.d_cidx); cudaFree(r.d_sidx);
    cudaFree(r.d_clag); cudaFree(r.d_slag);
    cudaFree(r.d_U); cudaFree(r.d_list); cudaFree(r.d_L); cudaFree(r.d_W);
    cudaFree(r.d_gmax); cudaFree(r.d_hkeys); cudaFree(r.d_gidx);
    cudaFree(r.d_flag); cudaFree(r.d_counters); cudaFree(r.d_sel); cudaFree(r.d_cut);
    cudaFree(r.d_ref); cudaFree(r.Xt); cudaFree(r.twM); cudaFree(r.twn);
    cudaFree(r.twp_f); cudaFree(r.twi_f); cudaFree(r.twide_f); cudaFree(r.swtw); cudaFree(r.sw_f); cudaFree(r.sx_f); cudaFree(r.sb_f); cudaFree(r.d_mid);
    for (int i = 0; i < 4; i++) if (r.ev[i]) cudaEventDestroy(r.ev[i]);
    if (r.aux) cudaStreamDestroy(r.aux);
    if (r.h_pin) cudaFreeHost(r.h_pin);
    memset(&r, 0, sizeof(r));
}

extern "C" void muse_ctx_destroy(muse_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (RunScratch &r : c->pool) scratch_free(r);
    for (int i = 0; i < 3; i++) {
        if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    cudaFree(c->tc_a); cudaFree(c->tc_b); cudaFree(c->tc_mid); cudaFree(c->tc_amid); cudaFree(c->tc_ptrs);
    cudaFree(c->d_multi_q);
    cudaFree(c->d_multi_next);
    for (int i = 0; i < 5; i++)
        if (c->mt_ev[i]) cudaEventDestroy(c->mt_ev[i]);
    cudaFree(c->d_mt);
    if (c->h_mt) cudaFreeHost(c->h_mt);
    cudaFree(c->d_multi_refs);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int muse_ctx_synchronize(muse_ctx *c) {
    if (!c) return fail(MUSE_ERR_INVALID_ARG, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return MUSE_OK;
}

extern "C" int muse_ctx_set_stream(muse_ctx *c, void *cuda_stream) {
    if (!c) return fail(MUSE_ERR_INVALID_ARG, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return MUSE_OK;
}

extern "C" int muse_host_alloc(void **out, int64_t bytes) {
    if (!out || bytes < 0) return fail(MUSE_ERR_INVALID_ARG, "muse_host_alloc: bad argument");
    CU(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return MUSE_OK;
}

extern "C" void muse_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------
// series store
// ------------------------------------------------------------------------------------
static int group_reserve(muse_group *g, int64_t want) {
    if (want <= g->cap) return MUSE_OK;
    int64_t ncap = std::max<int64_t>(want, g->cap + g->cap / 2);
    ncap = std::max<int64_t>(ncap, 16);
    double *nslab = nullptr;
    int32_t *nlab = nullptr;
    RowStat *nstat = nullptr;
    auto bail = [&](int code) {
        cudaFree(nslab);
        cudaFree(nlab);
        cudaFree(nstat);
        return code;
    };
    cudaError_t e = cudaMalloc(&nslab, sizeof(double) * (size_t)ncap * (size_t)g->ld);
    if (e == cudaSuccess && g->nkeys > 0) e = cudaMalloc(&nlab, sizeof(int32_t) * (size_t)ncap * (size_t)g->nkeys);
    if (e == cudaSuccess) e = cudaMalloc(&nstat, sizeof(RowStat) * (size_t)ncap);
    if (e == cudaSuccess && g->size > 0) {
        e = cudaMemcpyAsync(nslab, g->slab, sizeof(double) * (size_t)g->size * (size_t)g->ld, cudaMemcpyDeviceToDevice, g->ctx->stream);
        for (int k = 0; k < g->nkeys && e == cudaSuccess; k++)
            e = cudaMemcpyAsync(nlab + (size_t)k * ncap, g->labels + (size_t)k * g->cap, sizeof(int32_t) * (size_t)g->size,
                                cudaMemcpyDeviceToDevice, g->ctx->stream);
        if (e == cudaSuccess && g->stats_upto > 0)
            e = cudaMemcpyAsync(nstat, g->row_stat, sizeof(RowStat) * (size_t)g->stats_upto, cudaMemcpyDeviceToDevice, g->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->ctx->stream);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return bail(fail(e == cudaErrorMemoryAllocation ? MUSE_ERR_OUT_OF_MEMORY : MUSE_ERR_CUDA, "growing the store to %lld series: %s",
                         (long long)ncap, cudaGetErrorString(e)));
    }
    if (g->slab) cudaFree(g->slab);
    if (g->labels) cudaFree(g->labels);
    if (g->row_stat) cudaFree(g->row_stat);
    g->slab = nslab;
    g->labels = nlab;
    g->row_stat = nstat;
    g->cap = ncap;
    g->stats_cap = ncap;
    return MUSE_OK;
}

// RowStat of rows [first, first + count): the batched mean / sample-std kernel of the z-normalisation (xcorr.go:84-95),
// queued right behind the copy that brought the rows in, so that no Run ever pays a separate pass over the slab.
static int queue_row_stats(muse_group *g, int64_t first, int64_t count) {
    if (count <= 0) return MUSE_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((count + 7) / 8, (int64_t)g->ctx->sm_count * 16);
    row_stats_kernel<<<grid, 256, 0, g->ctx->stream>>>(g->slab, g->ld, (int)g->N, first, count, g->row_stat);
    CU(cudaGetLastError());
    return MUSE_OK;
}

extern "C" int muse_group_create(muse_ctx *ctx, int64_t series_len, int32_t n_label_keys, int64_t capacity_hint,
                                 muse_group **out) {
    if (!ctx || !out) return fail(MUSE_ERR_INVALID_ARG, "muse_group_create: NULL argument");
    if (series_len < 2)
        return fail(MUSE_ERR_INVALID_ARG, "series length %lld: the sample std needs at least 2 samples (xcorr.go:88)",
                    (long long)series_len);
    if (n_label_keys < 0 || n_label_keys > MUSE_MAX_LABEL_KEYS)
        return fail(MUSE_ERR_INVALID_ARG, "n_label_keys %d not in [0,%d]", n_label_keys, MUSE_MAX_LABEL_KEYS);
    CU(cudaSetDevice(ctx->device));
    muse_group *g = new muse_group();
    memset(g, 0, sizeof(*g));
    g->ctx = ctx;
    g->N = series_len;
    g->ld = (series_len + 15) / 16 * 16;
    g->nkeys = n_label_keys;
    for (int k = 0; k < MUSE_MAX_LABEL_KEYS; k++) g->max_id[k] = -1;
    int rc = group_reserve(g, std::max<int64_t>(capacity_hint, 16));
    if (rc != MUSE_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return MUSE_OK;
}

extern "C" void muse_group_destroy(muse_group *g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    if (g->slab) cudaFree(g->slab);
    if (g->labels) cudaFree(g->labels);
    if (g->row_stat) cudaFree(g->row_stat);
    delete g;
}

// Pageable host rows -> slab through a ring of three pinned buffers: MUSE_STAGE_THREADS host threads copy chunk c into
// buffer c % 3 while chunk c - 1 is on the wire (at most two DMA copies outstanding, so a buffer is free again when its
// turn comes).  The chunk size is a whole number of rows.
#define MUSE_STAGE_BYTES ((size_t)64 << 20)
static int append_staged(muse_group *g, const double *rows, int64_t n_series) {
    muse_ctx *c = g->ctx;
    cudaStream_t st = c->stream;
    for (int i = 0; i < 3; i++) {
        if (!c->h_stage[i]) CU(cudaHostAlloc((void **)&c->h_stage[i], MUSE_STAGE_BYTES, cudaHostAllocDefault));
        if (!c->stage_ev[i]) CU(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    const size_t row_bytes = sizeof(double) * (size_t)g->N;
    if (row_bytes > MUSE_STAGE_BYTES) return fail(MUSE_ERR_UNSUPPORTED, "a row of %zu bytes exceeds the staging buffer", row_bytes);
    const int64_t rows_per_chunk = (int64_t)(MUSE_STAGE_BYTES / row_bytes);
    const int64_t n_chunks = (n_series + rows_per_chunk - 1) / rows_per_chunk;
    unsigned hw = std::thread::hardware_concurrency();
    const char *env = getenv("MUSE_STAGE_THREADS");
    int n_threads = env ? atoi(env) : (int)std::min<unsigned>(hw ? hw : 4u, 12u);
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_series * row_bytes < ((size_t)8 << 20)) n_threads = 1;      // small appends: not worth the threads
    std::vector<std::atomic<int>> done((size_t)n_chunks);
    for (auto &d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<int64_t> writable(std::min<int64_t>(n_chunks, 2));      // chunks whose buffer may be filled
    std::atomic<bool> abort_flag(false);
    auto worker = [&](int w) {
        for (int64_t ch = 0; ch < n_chunks; ch++) {
            if (abort_flag.load(std::memory_order_relaxed)) return;
            while (writable.load(std::memory_order_acquire) <= ch) {
                if (abort_flag.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            const int64_t r0 = ch * rows_per_chunk, nr = std::min<int64_t>(rows_per_chunk, n_series - r0);
            const size_t bytes = (size_t)nr * row_bytes;
            const size_t lo = bytes * (size_t)w / (size_t)n_threads / 64 * 64, hi = (w + 1 == n_threads) ? bytes : bytes * (size_t)(w + 1) / (size_t)n_threads / 64 * 64;
            if (hi > lo) memcpy(c->h_stage[ch % 3] + lo, reinterpret_cast<const unsigned char *>(rows) + (size_t)r0 * row_bytes + lo, hi - lo);
            done[(size_t)ch].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < n_threads; w++) pool.emplace_back(worker, w);
    cudaError_t e = cudaSuccess;
    // the calling thread is worker 0 and the one that talks to the device
    for (int64_t ch = 0; ch < n_chunks && e == cudaSuccess; ch++) {
        const int64_t r0 = ch * rows_per_chunk, nr = std::min<int64_t>(rows_per_chunk, n_series - r0);
        {
            const size_t bytes = (size_t)nr * row_bytes;
            const size_t hi = (n_threads == 1) ? bytes : bytes / (size_t)n_threads / 64 * 64;
            if (hi > 0) memcpy(c->h_stage[ch % 3], reinterpret_cast<const unsigned char *>(rows) + (size_t)r0 * row_bytes, hi);
            done[(size_t)ch].fetch_add(1, std::memory_order_release);
        }
        while (done[(size_t)ch].load(std::memory_order_acquire) < n_threads) std::this_thread::yield();
        double *dst = g->slab + (size_t)(g->size + r0) * g->ld;
        if (g->ld == g->N) e = cudaMemcpyAsync(dst, c->h_stage[ch % 3], (size_t)nr * row_bytes, cudaMemcpyHostToDevice, st);
        else e = cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)g->ld, c->h_stage[ch % 3], row_bytes, row_bytes, (size_t)nr, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(c->stage_ev[ch % 3], st);
        if (e == cudaSuccess && ch >= 1) e = cudaEventSynchronize(c->stage_ev[(ch - 1) % 3]);      // buffer (ch + 2) % 3 is free again
        writable.store(ch + 3, std::memory_order_release);
    }
    if (e != cudaSuccess) abort_flag.store(true);
    writable.store(n_chunks + 3, std::memory_order_release);
    for (auto &t : pool) t.join();
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(MUSE_ERR_CUDA, "staged host->device copy failed: %s", cudaGetErrorString(e));
    }
    return MUSE_OK;
}

extern "C" int muse_group_append(muse_group *g, const double *rows, int64_t n_series, int64_t series_len,
                                 const int32_t *label_ids) {
    if (!g || (!rows && n_series > 0)) return fail(MUSE_ERR_INVALID_ARG, "muse_group_append: NULL argument");
    if (n_series < 0) return fail(MUSE_ERR_INVALID_ARG, "n_series < 0");
    if (series_len != g->N)   // group.go:45-51
        return fail(MUSE_ERR_LENGTH_MISMATCH, "Timeseries has length %lld, but current group has length %lld",
                    (long long)series_len, (long long)g->N);
    if (g->nkeys > 0 && !label_ids && n_series > 0) return fail(MUSE_ERR_INVALID_ARG, "label_ids is NULL but the group has %d label keys", g->nkeys);
    if (n_series == 0) return MUSE_OK;
    if (g->size + n_series > 0x7fffffffLL) return fail(MUSE_ERR_UNSUPPORTED, "more than 2^31-1 series in one store");
    CU(cudaSetDevice(g->ctx->device));
    int rc = group_reserve(g, g->size + n_series);
    if (rc != MUSE_OK) return rc;
    cudaStream_t st = g->ctx->stream;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, rows) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
    if (!pinned) {
        // pageable rows (a Go slice, a numpy array): a direct cudaMemcpyAsync stages them through the driver's single
        // bounce buffer at a fraction of the link rate -> chunks are copied into a pinned ring by several host threads
        // while the previous chunk is on the wire
        rc = append_staged(g, rows, n_series);
        if (rc != MUSE_OK) return rc;
    } else if (g->ld == g->N) {   // rows are already at the slab's pitch: one flat copy (a pitched copy of 10^6 rows is not always at DMA rate)
        CU(cudaMemcpyAsync(g->slab + (size_t)g->size * g->ld, rows, sizeof(double) * (size_t)g->N * (size_t)n_series,
                           cudaMemcpyHostToDevice, st));
    } else {
        CU(cudaMemcpy2DAsync(g->slab + (size_t)g->size * g->ld, sizeof(double) * (size_t)g->ld, rows,
                             sizeof(double) * (size_t)g->N, sizeof(double) * (size_t)g->N, (size_t)n_series,
                             cudaMemcpyHostToDevice, st));
    }
    if (g->nkeys > 0) {
        // host [n_series][nkeys] -> device SoA [nkeys][cap]
        std::vector<int32_t> col((size_t)n_series);
        for (int k = 0; k < g->nkeys; k++) {
            int32_t mx = g->max_id[k];
            for (int64_t i = 0; i < n_series; i++) {
                int32_t v = label_ids[(size_t)i * g->nkeys + k];
                if (v < -1) v = -1;
                col[(size_t)i] = v;
                mx = std::max(mx, v);
            }
            g->max_id[k] = mx;
            CU(cudaMemcpyAsync(g->labels + (size_t)k * g->cap + g->size, col.data(), sizeof(int32_t) * (size_t)n_series,
                               cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));   // col is reused
        }
    }
    rc = queue_row_stats(g, g->size, n_series);
    if (rc != MUSE_OK) return rc;
    CU(cudaStreamSynchronize(st));
    g->size += n_series;
    g->stats_upto = g->size;
    return MUSE_OK;
}

__global__ void max_id_kernel(const int32_t *ids, int64_t n, int32_t *out) {
    int32_t m = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, ids[i]);
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void transpose_labels_kernel(const int32_t *src, int64_t n, int nkeys, int32_t *dst, int64_t cap, int64_t off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < nkeys; k++) dst[(size_t)k * cap + off + i] = max(src[(size_t)i * nkeys + k], -1);
}

static int refresh_max_ids(muse_group *g, int64_t off, int64_t n) {
    if (g->nkeys == 0 || n == 0) return MUSE_OK;
    int32_t *d = nullptr;
    CU(cudaMalloc(&d, sizeof(int32_t) * MUSE_MAX_LABEL_KEYS));
    CU(cudaMemsetAsync(d, 0xff, sizeof(int32_t) * MUSE_MAX_LABEL_KEYS, g->ctx->stream));
    for (int k = 0; k < g->nkeys; k++)
        max_id_kernel<<<256, 256, 0, g->ctx->stream>>>(g->labels + (size_t)k * g->cap + off, n, d + k);
    int32_t h[MUSE_MAX_LABEL_KEYS];
    CU(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g->ctx->stream));
    CU(cudaStreamSynchronize(g->ctx->stream));
    cudaFree(d);
    for (int k = 0; k < g->nkeys; k++) g->max_id[k] = std::max(g->max_id[k], h[k]);
    return MUSE_OK;
}

extern "C" int muse_group_append_device(muse_group *g, const double *d_rows, int64_t n_series, int64_t series_len,
                                        const int32_t *d_label_ids) {
    if (!g || (!d_rows && n_series > 0)) return fail(MUSE_ERR_INVALID_ARG, "muse_group_append_device: NULL argument");
    if (series_len != g->N)
        return fail(MUSE_ERR_LENGTH_MISMATCH, "Timeseries has length %lld, but current group has length %lld",
                    (long long)series_len, (long long)g->N);
    if (g->nkeys > 0 && !d_label_ids && n_series > 0) return fail(MUSE_ERR_INVALID_ARG, "d_label_ids is NULL");
    if (n_series <= 0) return n_series == 0 ? MUSE_OK : fail(MUSE_ERR_INVALID_ARG, "n_series < 0");
    if (g->size + n_series > 0x7fffffffLL) return fail(MUSE_ERR_UNSUPPORTED, "more than 2^31-1 series in one store");
    CU(cudaSetDevice(g->ctx->device));
    int rc = group_reserve(g, g->size + n_series);
    if (rc != MUSE_OK) return rc;
    cudaStream_t st = g->ctx->stream;
    CU(cudaMemcpy2DAsync(g->slab + (size_t)g->size * g->ld, sizeof(double) * (size_t)g->ld, d_rows,
                         sizeof(double) * (size_t)g->N, sizeof(double) * (size_t)g->N, (size_t)n_series,
                         cudaMemcpyDeviceToDevice, st));
    if (g->nkeys > 0) {
        transpose_labels_kernel<<<(unsigned)((n_series + 255) / 256), 256, 0, st>>>(d_label_ids, n_series, g->nkeys, g->labels,
                                                                                  g->cap, g->size);
        CU(cudaGetLastError());
        rc = refresh_max_ids(g, g->size, n_series);
        if (rc != MUSE_OK) return rc;
    }
    rc = queue_row_stats(g, g->size, n_series);
    if (rc != MUSE_OK) return rc;
    CU(cudaStreamSynchronize(st));
    g->size += n_series;
    g->stats_upto = g->size;
    return MUSE_OK;
}

// one block per series row; coalesced 8-byte stores.  The row's RowStat (row_stats_kernel's sums about the first
// sample) is accumulated while the samples are generated: a synthetic store never needs a pass of its own for them.
__global__ void __launch_bounds__(256)
synth_rows_kernel(double *slab, int64_t ld, int64_t N, int64_t row0, int64_t n_series, uint64_t seed,
                  int64_t first_index, int32_t *lab0, int32_t *lab1, RowStat *stat, int variant) {
    __shared__ double red[2][8];
    for (int64_t r = blockIdx.x; r < n_series; r += gridDim.x) {
        const int64_t gi = first_index + r;
        const SynthSeries sp = synth_params(seed, gi, N, variant);
        double *row = slab + (size_t)(row0 + r) * ld;
        const double pivot = synth_value(seed, gi, sp, 0);
        double s1 = 0.0, s2 = 0.0;
        for (int64_t t = threadIdx.x; t < ld; t += blockDim.x) {
            const double v = t < N ? synth_value(seed, gi, sp, t) : 0.0;
            row[t] = v;
            if (t < N) {
                const double x = v - pivot;
                s1 += x;
                s2 = fma(x, x, s2);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = s1;
            red[1][threadIdx.x >> 5] = s2;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s1 = s2 = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
                s1 += red[0][w];
                s2 += red[1][w];
            }
            const RowStat rsr = make_row_stat(pivot, s1, s2, (int)N);
            stat[row0 + r] = rsr;
            if (N & 1) row[N] = rsr.mean;      // odd length: the pad column the screening kernels read with the last sample (RowStat)
            if (lab0) lab0[row0 + r] = (int32_t)(gi / 1000);
            if (lab1) lab1[row0 + r] = (int32_t)(gi % 1000);
        }
        __syncthreads();
    }
}

extern "C" int muse_group_append_synthetic_ex(muse_group *g, int64_t n_series, uint64_t seed, int64_t first_index, int32_t variant);
extern "C" int muse_group_append_synthetic(muse_group *g, int64_t n_series, uint64_t seed, int64_t first_index) {
    return muse_group_append_synthetic_ex(g, n_series, seed, first_index, 0);
}

extern "C" int muse_group_append_synthetic_ex(muse_group *g, int64_t n_series, uint64_t seed, int64_t first_index, int32_t variant) {
    if (!g || n_series < 0 || variant < 0 || variant > 1) return fail(MUSE_ERR_INVALID_ARG, "muse_group_append_synthetic: bad argument");
    if (n_series == 0) return MUSE_OK;
    if (g->size + n_series > 0x7fffffffLL) return fail(MUSE_ERR_UNSUPPORTED, "more than 2^31-1 series in one store");
    CU(cudaSetDevice(g->ctx->device));
    int rc = group_reserve(g, g->size + n_series);
    if (rc != MUSE_OK) return rc;
    int32_t *l0 = g->nkeys >= 1 ? g->labels : nullptr;
    int32_t *l1 = g->nkeys >= 2 ? g->labels + (size_t)g->cap : nullptr;
    if (g->nkeys > 2)
        CU(cudaMemsetAsync(g->labels + (size_t)2 * g->cap, 0, sizeof(int32_t) * (size_t)g->cap * (size_t)(g->nkeys - 2), g->ctx->stream));
    const unsigned grid = (unsigned)std::min<int64_t>(n_series, (int64_t)g->ctx->sm_count * 16);
    synth_rows_kernel<<<grid, 256, 0, g->ctx->stream>>>(g->slab, g->ld, g->N, g->size, n_series, seed, first_index, l0, l1, g->row_stat, variant);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(g->ctx->stream));
    g->stats_upto = g->size + n_series;
    if (g->nkeys >= 1) g->max_id[0] = std::max<int32_t>(g->max_id[0], (int32_t)((first_index + n_series - 1) / 1000));
    if (g->nkeys >= 2) g->max_id[1] = std::max<int32_t>(g->max_id[1], (int32_t)std::min<int64_t>(999, first_index + n_series - 1));
    for (int k = 2; k < g->nkeys; k++) g->max_id[k] = std::max(g->max_id[k], 0);
    g->size += n_series;
    return MUSE_OK;
}

// label id of key k for global series index i: (i / div[k]) % mod[k]
__global__ void synth_labels_kernel(int32_t *labels, int64_t cap, int nkeys, int64_t row0, int64_t n_series, int64_t first_index,
                                    const int64_t d0, const int64_t m0, const int64_t d1, const int64_t m1, const int64_t d2,
                                    const int64_t m2, const int64_t d3, const int64_t m3) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_series) return;
    const int64_t gi = first_index + r;
    const int64_t d[4] = {d0, d1, d2, d3}, m[4] = {m0, m1, m2, m3};
    for (int k = 0; k < nkeys && k < 4; k++) labels[(size_t)k * cap + row0 + r] = (int32_t)((gi / d[k]) % m[k]);
}

extern "C" int muse_group_set_synthetic_labels(muse_group *g, const int64_t *div, const int64_t *mod) {
    if (!g || !div || !mod) return fail(MUSE_ERR_INVALID_ARG, "muse_group_set_synthetic_labels: NULL argument");
    if (g->nkeys > 4) return fail(MUSE_ERR_UNSUPPORTED, "synthetic labels for at most 4 keys");
    if (g->size == 0 || g->nkeys == 0) return MUSE_OK;
    int64_t d[4] = {1, 1, 1, 1}, m[4] = {1, 1, 1, 1};
    for (int k = 0; k < g->nkeys; k++) {
        if (div[k] < 1 || mod[k] < 1 || mod[k] > 0x7fffffffLL) return fail(MUSE_ERR_INVALID_ARG, "label divisor / modulus out of range");
        d[k] = div[k];
        m[k] = mod[k];
    }
    CU(cudaSetDevice(g->ctx->device));
    synth_labels_kernel<<<(unsigned)((g->size + 255) / 256), 256, 0, g->ctx->stream>>>(g->labels, g->cap, g->nkeys, 0, g->size,
                                                                                       g->global_offset, d[0], m[0], d[1], m[1],
                                                                                       d[2], m[2], d[3], m[3]);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(g->ctx->stream));
    for (int k = 0; k < g->nkeys; k++)      // an upper bound is all the group table needs
        g->max_id[k] = (int32_t)std::min<int64_t>(m[k] - 1, (g->global_offset + g->size - 1) / d[k]);
    return MUSE_OK;
}

extern "C" void muse_synth_row(uint64_t seed, int64_t index, int64_t series_len, double *out_row) {
    const SynthSeries sp = synth_params(seed, index, series_len);
    for (int64_t t = 0; t < series_len; t++) out_row[t] = synth_value(seed, index, sp, t);
}

extern "C" void muse_synth_reference(uint64_t seed, int64_t series_len, double *out_row) {
    for (int64_t t = 0; t < series_len; t++) out_row[t] = synth_ref_value(seed, series_len, t);
}

extern "C" int64_t muse_group_size(const muse_group *g) { return g ? g->size : 0; }
extern "C" int64_t muse_group_series_len(const muse_group *g) { return g ? g->N : 0; }

extern "C" int muse_group_set_global_offset(muse_group *g, int64_t first_global_index) {
    if (!g || first_global_index < 0) return fail(MUSE_ERR_INVALID_ARG, "muse_group_set_global_offset: bad argument");
    g->global_offset = first_global_index;
    return MUSE_OK;
}

extern "C" int muse_group_clear(muse_group *g) {
    if (!g) return fail(MUSE_ERR_INVALID_ARG, "group is NULL");
    CU(cudaSetDevice(g->ctx->device));
    CU(cudaStreamSynchronize(g->ctx->stream));
    g->size = 0;
    g->stats_upto = 0;
    for (int k = 0; k < MUSE_MAX_LABEL_KEYS; k++) g->max_id[k] = -1;
    return MUSE_OK;
}

extern "C" int muse_group_read_rows(muse_group *g, int64_t first, int64_t n_rows, double *out_rows) {
    if (!g || !out_rows || first < 0 || n_rows < 0 || first + n_rows > g->size) return fail(MUSE_ERR_INVALID_ARG, "muse_group_read_rows: bad argument");
    if (n_rows == 0) return MUSE_OK;
    CU(cudaSetDevice(g->ctx->device));
    CU(cudaMemcpy2DAsync(out_rows, sizeof(double) * (size_t)g->N, g->slab + (size_t)first * g->ld, sizeof(double) * (size_t)g->ld,
                         sizeof(double) * (size_t)g->N, (size_t)n_rows, cudaMemcpyDeviceToHost, g->ctx->stream));
    CU(cudaStreamSynchronize(g->ctx->stream));
    return MUSE_OK;
}

extern "C" int muse_group_read_row(muse_group *g, int64_t local_index, double *out_row) {
    if (!g || !out_row || local_index < 0 || local_index >= g->size) return fail(MUSE_ERR_INVALID_ARG, "muse_group_read_row: bad argument");
    CU(cudaSetDevice(g->ctx->device));
    CU(cudaMemcpyAsync(out_row, g->slab + (size_t)local_index * g->ld, sizeof(double) * (size_t)g->N, cudaMemcpyDeviceToHost,
                       g->ctx->stream));
    CU(cudaStreamSynchronize(g->ctx->stream));
    return MUSE_OK;
}

// ------------------------------------------------------------------------------------
// batch
// ------------------------------------------------------------------------------------
static int64_t next_pow2(int64_t v) {   // == nextPowOf2 (xcorr.go:19-24) for 1 <= v < 2^29
    int64_t n = 1;
    while (n < v) n <<= 1;
    return n;
}

// fp32 tables of the screening kernel; leaves screen_ok = 0 when the shape has no screening kernel
// n = 128 .. 16384: kernels with the fused fp32 second stage (several series per warp up to 1024, one warp per series at 2048,
// one block per series above)
static bool block_small();
static int screen_is_fused(int log2m) { return log2m >= (block_small() ? 8 : 6) && log2m <= 13; }
static int screen_is_big(int log2m) { return log2m >= 11 && log2m <= 13; }      // muse_screen_big.cuh
static int screen_log2m_supported(int log2m) { return screen_is_fused(log2m); }
// n = 128 .. 1024: several series per warp (muse_screen_sub.cuh); MUSE_BLOCK_SMALL=1 keeps round 1's block kernel for A/B runs
static bool block_small() {
    static const bool v = getenv("MUSE_BLOCK_SMALL") != nullptr;
    return v;
}
static int screen_log2p(int log2m) { return (log2m >= 10 || !block_small()) ? 5 : (log2m == 9 ? 4 : 3); }

// The per-reference tables of the fp32 kernels, from Xt on the device: weights of the bound
// A[k] = |Xt[k]| * (1 at DC and Nyquist, else 2) rounded UP (the bound must not shrink), one 16-byte entry per
// mirror pair (k, M-k), k < M/2, with the split twiddle exp(-2*pi*i*k/n); the two reference coefficients of the
// second stage; and A[M/2], Xt[M/2] for the host (kernel parameters).
__global__ void screen_tables_kernel(const cd *__restrict__ Xt, int M, const float2 *__restrict__ swtw, float4 *__restrict__ sw,
                                     float4 *__restrict__ sx, float *__restrict__ sb, float *__restrict__ mid,
                                     const int32_t *__restrict__ std_zero) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M / 2) return;
    if (k == 0) mid[0] = __int_as_float(*std_zero);      // the reference's std-zero flag rides along: one copy to the host
    const cd a = Xt[k], c = Xt[M - k];
    if (k == M / 2) {
        mid[1] = __double2float_ru(hypot(a.x, a.y) * 2.0);
        mid[2] = (float)a.x;
        mid[3] = (float)a.y;
        return;
    }
    const float wa = __double2float_ru(hypot(a.x, a.y) * (k == 0 ? 1.0 : 2.0));
    const float wc = __double2float_ru(hypot(c.x, c.y) * (k == 0 ? 1.0 : 2.0));     // k = 0: M - k is the Nyquist bin
    const float2 w = swtw[k];
    sw[k] = make_float4(w.x, w.y, wa, wc);
    sx[k] = make_float4((float)a.x, (float)a.y, (float)c.x, (float)c.y);
    sb[k] = __double2float_ru(sqrt(2.0 * ((double)wa * (double)wa + (double)wc * (double)wc)) * (1.0 + 1e-15));
}

// Twiddle tables of FFT length n (they depend on n alone): filled when the scratch set was last used for another n.
static int ensure_ref_tables(muse_batch *b, int64_t ld) {
    const int64_t n = b->n, M = n / 2;
    cudaStream_t st = bstream(b);
    if (b->d_ref_cap < ld) {
        cudaFree(b->d_ref);
        b->d_ref = nullptr;
        CU(cudaMalloc(&b->d_ref, sizeof(double) * (size_t)ld));
        b->d_ref_cap = ld;
    }
    if (!b->d_mid) CU(cudaMalloc(&b->d_mid, sizeof(float) * 4));
    const bool want_screen = screen_log2m_supported(b->log2m);
    if (b->tab_n == n && (b->tab_screen || !want_screen)) return MUSE_OK;
    if (b->is_long) {      // the reference's transform has n entries, the twiddles are generated on the device
        cudaFree(b->Xt); cudaFree(b->twM); cudaFree(b->twn); cudaFree(b->twp_f); cudaFree(b->twi_f); cudaFree(b->twide_f); cudaFree(b->swtw); cudaFree(b->sw_f); cudaFree(b->sx_f); cudaFree(b->sb_f);
        b->Xt = b->twM = b->twn = nullptr;
        b->twp_f = b->twi_f = b->twide_f = nullptr; b->swtw = nullptr; b->sw_f = b->sx_f = nullptr; b->sb_f = nullptr;
        b->tab_n = 0;
        b->tab_screen = 0;
        CU(cudaMalloc(&b->Xt, sizeof(cd) * (size_t)n));
        CU(cudaMalloc(&b->twM, sizeof(cd) * (size_t)n));
        CU(launch_long_twiddles(b->twM, b->log2m + 1, st));
        b->tab_n = n;
        return MUSE_OK;
    }
    cudaFree(b->Xt); cudaFree(b->twM); cudaFree(b->twn); cudaFree(b->twp_f); cudaFree(b->twi_f); cudaFree(b->twide_f); cudaFree(b->swtw); cudaFree(b->sw_f); cudaFree(b->sx_f); cudaFree(b->sb_f);
    b->Xt = b->twM = b->twn = nullptr;
    b->twp_f = b->twi_f = b->twide_f = nullptr; b->swtw = nullptr; b->sw_f = b->sx_f = nullptr; b->sb_f = nullptr;
    b->tab_n = 0;
    b->tab_screen = 0;
    const long double PI2 = 6.283185307179586476925286766559005768L;
    const int log2p = exact_log2p(b->log2m);
    std::vector<cd> twM((size_t)M + 16);
    fill_pass_twiddles(b->log2m, log2p, twM.data(), [&](long long num, long long den) {
        return cd{(double)cosl(-PI2 * num / den), (double)sinl(-PI2 * num / den)};
    });
    // twiddle tables, correctly rounded from long double
    std::vector<cd> twn((size_t)(M / 2 + 1));
    for (int64_t k = 0; k <= M / 2; k++) twn[(size_t)k] = cd{(double)cosl(-PI2 * k / n), (double)sinl(-PI2 * k / n)};
    CU(cudaMalloc(&b->Xt, sizeof(cd) * (size_t)(M + 1)));
    CU(cudaMalloc(&b->twM, sizeof(cd) * twM.size()));
    CU(cudaMalloc(&b->twn, sizeof(cd) * (size_t)(M / 2 + 1)));
    CU(cudaMemcpyAsync(b->twM, twM.data(), sizeof(cd) * twM.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b->twn, twn.data(), sizeof(cd) * (size_t)(M / 2 + 1), cudaMemcpyHostToDevice, st));
    std::vector<cf> twp, twi, twide;
    std::vector<float2> swtw;
    if (want_screen) {
        if (screen_is_big(b->log2m)) {
            twi.resize((size_t)10 * (M / 32) + 31 * (M / 1024) + 1024 + 31 * 32);      // ScreenBigCfg::TW_TOTAL
            fill_big_twiddles(b->log2m, twi.data(), [&](long long num, long long den) {
                return cf{(float)cosl(-PI2 * num / den), (float)sinl(-PI2 * num / den)};
            });
            CU(cudaMalloc(&b->twi_f, sizeof(cf) * twi.size()));
            CU(cudaMemcpyAsync(b->twi_f, twi.data(), sizeof(cf) * twi.size(), cudaMemcpyHostToDevice, st));
            if (b->log2m == 13 && getenv("MUSE_WIDE13")) {
                twide.resize(ScreenWideCfg::TW_TOTAL);
                fill_wide_twiddles(twide.data(), [&](long long num, long long den) {
                    return cf{(float)cosl(-PI2 * num / den), (float)sinl(-PI2 * num / den)};
                });
                CU(cudaMalloc(&b->twide_f, sizeof(cf) * twide.size()));
                CU(cudaMemcpyAsync(b->twide_f, twide.data(), sizeof(cf) * twide.size(), cudaMemcpyHostToDevice, st));
            }
        }
        twp.resize((size_t)M + 64);
        fill_pass_twiddles(b->log2m, screen_log2p(b->log2m), twp.data(), [&](long long num, long long den) {
            return cf{(float)cosl(-PI2 * num / den), (float)sinl(-PI2 * num / den)};
        });
        swtw.resize((size_t)M / 2);
        for (int64_t k = 0; k < M / 2; k++) swtw[(size_t)k] = make_float2((float)cosl(-PI2 * k / n), (float)sinl(-PI2 * k / n));
        CU(cudaMalloc(&b->twp_f, sizeof(cf) * twp.size()));
        CU(cudaMalloc(&b->swtw, sizeof(float2) * swtw.size()));
        CU(cudaMalloc(&b->sw_f, sizeof(float4) * (size_t)(M / 2)));
        CU(cudaMalloc(&b->sx_f, sizeof(float4) * (size_t)(M / 2)));
        CU(cudaMalloc(&b->sb_f, sizeof(float) * (size_t)(M / 2)));
        CU(cudaMemcpyAsync(b->twp_f, twp.data(), sizeof(cf) * twp.size(), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(b->swtw, swtw.data(), sizeof(float2) * swtw.size(), cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st));      // the host vectors go out of scope
    b->tab_n = n;
    b->tab_screen = want_screen ? 1 : 0;
    return MUSE_OK;
}

// NewBatch in two halves, so that the references of a multi-query launch share ONE round trip: queue = everything
// up to the copies of the std-zero flag and the middle-bin values into the batch's pinned mailbox; finish (after the
// stream has been synchronised) = the error of muse_batch.go:38-41 (the batch is destroyed) or the by-value parameters.
// d_ref_row != NULL: the reference row is already on the device, zero-padded to the slab's row pitch (muse_multi_run
// uploads the references of a launch with one copy).
extern "C" void muse_batch_destroy(muse_batch *b);
static int ensure_long_work(muse_batch *b, int64_t count);
static LongParams long_params(const muse_batch *b, const double *slab, int64_t count, const int32_t *idx, int signed_scores);

// as CU, for code that owns a half-built batch `b`: it goes back to the pool before the error is returned
#define CUB(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            muse_batch_destroy(b);                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? MUSE_ERR_OUT_OF_MEMORY : MUSE_ERR_CUDA,  \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

// The batch object with its pooled scratch, the tables of its FFT length and its small allocations: nothing is launched.
static int batch_alloc(muse_ctx *ctx, muse_group *g, int64_t ref_len, muse_batch **out) {
    if (!ctx || !g || !out) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_create: NULL argument");
    if (g->ctx != ctx) return fail(MUSE_ERR_INVALID_ARG, "group belongs to another context");
    if (ref_len != g->N)   // muse_batch.go:24-28
        return fail(MUSE_ERR_LENGTH_MISMATCH, "comparison group series (length %lld) does not have the same length as the reference (%lld)",
                    (long long)g->N, (long long)ref_len);
    const int64_t n = next_pow2(ref_len);   // muse_batch.go:35
    if (n > MUSE_MAX_FFT_LEN) return fail(MUSE_ERR_UNSUPPORTED, "FFT length %lld > %d", (long long)n, MUSE_MAX_FFT_LEN);
    CU(cudaSetDevice(ctx->device));
    muse_batch *b = new muse_batch();
    memset(b, 0, sizeof(*b));
    b->is_long = n > MUSE_MAX_FUSED_FFT_LEN;
    b->ctx = ctx;
    b->g = g;
    b->N = ref_len;
    b->n = n;
    const int64_t M = n / 2;
    int l = 0;
    while ((1LL << l) < M) l++;
    b->log2m = l;
    {   // scratch of an earlier batch on this context, if there is one: a set whose tables were filled for this FFT
        // length is preferred, then the largest (it fits most stores)
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->pool.empty()) {
            size_t best = 0;
            auto better = [&](const RunScratch &x, const RunScratch &y) {
                if ((x.tab_n == n) != (y.tab_n == n)) return x.tab_n == n;
                return x.scratch_cap > y.scratch_cap;
            };
            for (size_t i = 1; i < ctx->pool.size(); i++)
                if (better(ctx->pool[i], ctx->pool[best])) best = i;
            static_cast<RunScratch &>(*b) = ctx->pool[best];
            ctx->pool.erase(ctx->pool.begin() + (long)best);
        }
    }
    auto bail = [&](int code) {
        muse_batch_destroy(b);
        return code;
    };
    for (int i = 0; i < 4; i++)
        if (!b->ev[i] && cudaEventCreate(&b->ev[i]) != cudaSuccess) return bail(fail(MUSE_ERR_CUDA, "cudaEventCreate failed"));
    int rc = ensure_ref_tables(b, g->ld);
    if (rc) return bail(rc);
    cudaError_t e = cudaSuccess;
    if (!b->d_flag) e = cudaMalloc(&b->d_flag, sizeof(int32_t));
    if (e == cudaSuccess && !b->d_counters) e = cudaMalloc(&b->d_counters, sizeof(unsigned long long) * 4);
    if (e == cudaSuccess && !b->d_sel) e = cudaMalloc(&b->d_sel, sizeof(SelectState));
    if (e == cudaSuccess && !b->d_cut) e = cudaMalloc(&b->d_cut, sizeof(unsigned) * (4 + MUSE_CUT_WORDS));
    if (e == cudaSuccess && !b->h_pin) {
        b->h_pin_bytes = (size_t)4 << 20;
        e = cudaHostAlloc((void **)&b->h_pin, b->h_pin_bytes, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return bail(fail(e == cudaErrorMemoryAllocation ? MUSE_ERR_OUT_OF_MEMORY : MUSE_ERR_CUDA, "muse_batch_create: %s", cudaGetErrorString(e)));
    }
    *out = b;
    return MUSE_OK;
}

static int batch_create_queue(muse_ctx *ctx, muse_group *g, const double *ref, int64_t ref_len, muse_batch **out,
                              const double *d_ref_row = nullptr) {
    if (!ref) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_create: NULL argument");
    muse_batch *b = nullptr;
    int rc = batch_alloc(ctx, g, ref_len, &b);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const int64_t M = b->n / 2;
    if (!d_ref_row) {      // (the select state is cleared where it is used)
        CUB(cudaMemsetAsync(b->d_ref, 0, sizeof(double) * (size_t)g->ld, st));
        CUB(cudaMemcpyAsync(b->d_ref, ref, sizeof(double) * (size_t)ref_len, cudaMemcpyHostToDevice, st));
    }
    // X on the device through the same forward path the series take
    ExactParams p;
    memset(&p, 0, sizeof(p));
    p.slab = d_ref_row ? d_ref_row : b->d_ref;
    p.ld = g->ld;
    p.count = 1;
    p.N = (int)ref_len;
    p.twM = b->twM;
    p.twn = b->twn;
    p.out_X = b->Xt;
    p.out_flag = b->d_flag;
    if (b->is_long) {
        int rcl = ensure_long_work(b, 1);
        if (rcl) {
            muse_batch_destroy(b);
            return rcl;
        }
        LongParams lp = long_params(b, p.slab, 1, nullptr, 0);
        lp.out_X = b->Xt;
        CUB(launch_long(MODE_REF, lp, b->d_long, b->long_pairs, st));
    } else {
        CUB(launch_exact(MODE_REF, b->log2m, p, st));
    }
    b->screen_ok = 0;
    const bool screen = b->tab_screen != 0;
    if (screen) {
        screen_tables_kernel<<<(unsigned)((M / 2 + 1 + 255) / 256), 256, 0, st>>>(b->Xt, (int)M, b->swtw, b->sw_f, b->sx_f, b->sb_f, b->d_mid, b->d_flag);
        CUB(cudaGetLastError());
    }
    // one round trip: the std-zero flag of the reference and the two middle-bin values the kernels take by value
    int32_t *h_flag = reinterpret_cast<int32_t *>(b->h_pin);
    float *h_mid = reinterpret_cast<float *>(b->h_pin + 16);
    if (screen) CUB(cudaMemcpyAsync(h_mid, b->d_mid, sizeof(float) * 4, cudaMemcpyDeviceToHost, st));      // [0] carries the flag
    else CUB(cudaMemcpyAsync(h_flag, b->d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    b->screen_ok = screen ? -1 : 0;      // -1: pending until batch_create_finish
    *out = b;
    return MUSE_OK;
}

static int batch_create_finish(muse_batch *b) {
    const float *h_mid = reinterpret_cast<const float *>(b->h_pin + 16);
    int32_t flag;
    memcpy(&flag, b->screen_ok == -1 ? static_cast<const void *>(h_mid) : static_cast<const void *>(b->h_pin), sizeof(flag));
    if (flag) {   // muse_batch.go:38-41
        muse_batch_destroy(b);
        return fail(MUSE_ERR_STDDEV_ZERO, "Invalid input query, Standard deviation of zero");
    }
    if (b->screen_ok == -1) {
        b->a_mid = h_mid[1];
        b->x_mid = cf{h_mid[2], h_mid[3]};
        b->screen_ok = 1;
    }
    return MUSE_OK;
}


extern "C" int muse_batch_create(muse_ctx *ctx, muse_group *g, const double *ref, int64_t ref_len, muse_batch **out) {
    muse_batch *b = nullptr;
    int rc = batch_create_queue(ctx, g, ref, ref_len, &b);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        muse_batch_destroy(b);
        return fail(MUSE_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    }
    rc = batch_create_finish(b);
    if (rc) return rc;
    *out = b;
    return MUSE_OK;
}

static void free_scratch(muse_batch *b) {
    cudaFree(b->d_score); cudaFree(b->d_lag); cudaFree(b->d_slot);
    cudaFree(b->d_ckey); cudaFree(b->d_skey); cudaFree(b->d_cidx); cudaFree(b->d_sidx);
    cudaFree(b->d_clag); cudaFree(b->d_slag);
    b->d_clag = b->d_slag = nullptr;
    cudaFree(b->d_U); cudaFree(b->d_list);
    b->d_U = nullptr; b->d_list = nullptr;
    b->d_score = nullptr; b->d_lag = nullptr; b->d_slot = nullptr;
    b->d_ckey = b->d_skey = nullptr; b->d_cidx = b->d_sidx = nullptr;
    b->scratch_cap = 0;
}

extern "C" void muse_batch_destroy(muse_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(bstream(b));
    {   // the store-sized scratch goes back to the context (at most MUSE_SCRATCH_POOL sets are kept: a multi-query launch
        // group has 256 batches alive at once)
        std::lock_guard<std::mutex> lk(b->ctx->mu);
        if (b->ctx->pool.size() < MUSE_SCRATCH_POOL) {
            b->ctx->pool.push_back(static_cast<RunScratch &>(*b));
            memset(static_cast<RunScratch *>(b), 0, sizeof(RunScratch));
        }
    }
    scratch_free(*b);
    cudaFree(b->d_long);
    delete b;
}

extern "C" int64_t muse_batch_fft_len(const muse_batch *b) { return b ? b->n : 0; }

static int ensure_scratch(muse_batch *b) {
    const int64_t S = b->g->size;
    if (S <= b->scratch_cap) return MUSE_OK;
    free_scratch(b);
    const int64_t cap = std::max<int64_t>(S, 1024);
    CU(cudaMalloc(&b->d_score, sizeof(double) * (size_t)cap));
    CU(cudaMalloc(&b->d_lag, sizeof(int32_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_slot, sizeof(int64_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_ckey, sizeof(unsigned long long) * (size_t)cap));
    CU(cudaMalloc(&b->d_skey, sizeof(unsigned long long) * (size_t)cap));
    CU(cudaMalloc(&b->d_cidx, sizeof(int32_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_sidx, sizeof(int32_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_clag, sizeof(int32_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_slag, sizeof(int32_t) * (size_t)cap));
    CU(cudaMalloc(&b->d_U, sizeof(float) * (size_t)cap));
    CU(cudaMalloc(&b->d_list, sizeof(int32_t) * (size_t)cap));
    b->scratch_cap = cap;
    return MUSE_OK;
}

static int check_batch(muse_batch *b) {
    if (!b) return fail(MUSE_ERR_INVALID_ARG, "batch is NULL");
    if (b->g->N != b->N)
        return fail(MUSE_ERR_LENGTH_MISMATCH, "comparison group length %lld != reference length %lld", (long long)b->g->N, (long long)b->N);
    return MUSE_OK;
}

// n > MUSE_MAX_FUSED_FFT_LEN: the work buffer of launch_long, sized for `count` series or 256 MB worth of pairs
static int ensure_long_work(muse_batch *b, int64_t count) {
    const int log2n = b->log2m + 1;
    const size_t per_pair = long_work_bytes(log2n, 1);
    // both buffers of a chunk together: MUSE_LONG_WORK_MB (default 256)
    static const size_t work_mb = getenv("MUSE_LONG_WORK_MB") ? (size_t)std::max(1, atoi(getenv("MUSE_LONG_WORK_MB"))) : 256;
    long long pairs = std::max<long long>(1, (long long)((work_mb << 20) / per_pair));
    pairs = std::min<long long>(pairs, std::max<long long>(1, (count + 1) / 2));
    if (b->d_long && b->long_pairs >= pairs) return MUSE_OK;
    cudaFree(b->d_long);
    b->d_long = nullptr;
    b->long_pairs = 0;
    b->d_long_bytes = long_work_bytes(log2n, pairs);
    CU(cudaMalloc(&b->d_long, b->d_long_bytes));
    b->long_pairs = pairs;
    return MUSE_OK;
}

static LongParams long_params(const muse_batch *b, const double *slab, int64_t count, const int32_t *idx, int signed_scores) {
    LongParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.slab = slab;
    lp.ld = b->g->ld;
    lp.count = count;
    lp.idx = idx;
    lp.N = (int)b->N;
    lp.log2n = b->log2m + 1;
    lp.signed_scores = signed_scores;
    lp.X = b->Xt;
    lp.tw = b->twM;
    lp.out_score = b->d_score;
    lp.out_lag = b->d_lag;
    lp.out_flag = b->d_flag;
    return lp;
}

// All series of the store through the exact kernel -> d_score / d_lag.
static int score_exact_all(muse_batch *b, int signed_scores, const int32_t *idx, int64_t count,
                           const unsigned long long *d_count = nullptr) {
    ExactParams p;
    memset(&p, 0, sizeof(p));
    p.count_ptr = d_count;
    p.slab = b->g->slab;
    p.ld = b->g->ld;
    p.count = count;
    p.idx = idx;
    p.N = (int)b->N;
    p.signed_scores = signed_scores;
    p.Xt = b->Xt;
    p.twM = b->twM;
    p.twn = b->twn;
    p.out_score = b->d_score;
    p.out_lag = b->d_lag;
    if (count > 0 && b->is_long) {
        if (d_count) return fail(MUSE_ERR_UNSUPPORTED, "device-side list lengths do not exist above FFT length %d", MUSE_MAX_FUSED_FFT_LEN);
        int rc = ensure_long_work(b, count);
        if (rc) return rc;
        LongParams lp = long_params(b, b->g->slab, count, idx, signed_scores);
        CU(launch_long(MODE_SCORE, lp, b->d_long, b->long_pairs, bstream(b)));
        b->timing.n_launches++;
    } else if (count > 0) {
        CU(launch_exact(MODE_SCORE, b->log2m, p, bstream(b)));
        b->timing.n_launches++;
    }
    return MUSE_OK;
}

extern "C" int muse_batch_score_all(muse_batch *b, int32_t signed_scores, double *scores, int32_t *lags) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!scores || !lags) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_score_all: NULL output");
    CU(cudaSetDevice(b->ctx->device));
    rc = ensure_scratch(b);
    if (rc) return rc;
    const int64_t S = b->g->size;
    rc = score_exact_all(b, signed_scores, nullptr, S);
    if (rc) return rc;
    CU(cudaMemcpyAsync(scores, b->d_score, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost, bstream(b)));
    CU(cudaMemcpyAsync(lags, b->d_lag, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToHost, bstream(b)));
    CU(cudaStreamSynchronize(bstream(b)));
    return MUSE_OK;
}

extern "C" int muse_batch_xcorr(muse_batch *b, int64_t local_index, double *cc, int32_t *std_zero) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!cc || local_index < 0 || local_index >= b->g->size) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_xcorr: bad argument");
    CU(cudaSetDevice(b->ctx->device));
    double *d_cc = nullptr;
    CU(cudaMalloc(&d_cc, sizeof(double) * (size_t)b->n));
    ExactParams p;
    memset(&p, 0, sizeof(p));
    p.slab = b->g->slab + (size_t)local_index * b->g->ld;
    p.ld = b->g->ld;
    p.count = 1;
    p.N = (int)b->N;
    p.Xt = b->Xt;
    p.twM = b->twM;
    p.twn = b->twn;
    p.out_score = d_cc;
    p.out_flag = b->d_flag;
    int32_t flag = 0;
    auto body = [&]() -> int {      // d_cc is freed on every path out of here
        if (b->is_long) {
            int rcl = ensure_long_work(b, 1);
            if (rcl) return rcl;
            LongParams lp = long_params(b, p.slab, 1, nullptr, 0);
            lp.out_score = d_cc;
            CU(launch_long(MODE_CC, lp, b->d_long, b->long_pairs, bstream(b)));
        } else {
            CU(launch_exact(MODE_CC, b->log2m, p, bstream(b)));
        }
        CU(cudaMemcpyAsync(cc, d_cc, sizeof(double) * (size_t)b->n, cudaMemcpyDeviceToHost, bstream(b)));
        CU(cudaMemcpyAsync(&flag, b->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, bstream(b)));
        CU(cudaStreamSynchronize(bstream(b)));
        return MUSE_OK;
    };
    rc = body();
    cudaFree(d_cc);
    if (rc) return rc;
    if (std_zero) *std_zero = flag;
    return MUSE_OK;
}

// ------------------------------------------------------------------------------------
// Run: scores -> group max -> filter -> top-N
// ------------------------------------------------------------------------------------
struct RunArgs {
    const int32_t *key_cols;
    int32_t n_key_cols;
    int64_t max_lag, top_n;
    double threshold;
    int32_t sign_filter, mode, signed_scores;
    int32_t list_only;   // fused runs: the caller reads scores only through the exact list (no NaN fill of the score array)
    int32_t group_cut;   // grouped fused runs whose top-N is final here (one GPU): members below the top_n-th certain group need no exact score
};

struct Rec {
    unsigned long long key;   // |score| bits
    int32_t idx;
    int32_t lagsgn;           // 2*lag + (score < 0)
    double score() const {
        double a;
        memcpy(&a, &key, sizeof(a));
        return (lagsgn & 1) ? -a : a;
    }
    int32_t lag() const { return (lagsgn - (lagsgn & 1)) / 2; }
};

// Key bits per group-by column.  64 / ncols each when every column's cardinality fits: that packing does not depend on the
// store, so the shards of a multi-GPU run merge on it.  Otherwise ceil(log2(cardinality)) of THIS store's columns, which
// any number of keys up to MUSE_MAX_KEY_COLS can use as long as the widths sum to at most 64 bits (*rank_independent = 0).
static int key_bits(const muse_group *g, const int32_t *key_cols, int n_key_cols, int *bits, int *rank_independent) {
    if (n_key_cols > MUSE_MAX_KEY_COLS) return fail(MUSE_ERR_UNSUPPORTED, "grouping by more than %d label keys", MUSE_MAX_KEY_COLS);
    const int fixed = 64 / n_key_cols;
    bool fits = true;
    int total = 0;
    for (int c = 0; c < n_key_cols; c++) {
        const int col = key_cols[c];
        if (col < 0 || col >= g->nkeys) return fail(MUSE_ERR_INVALID_ARG, "key column %d not in [0,%d)", col, g->nkeys);
        const int64_t card = (int64_t)g->max_id[col] + 2;      // ids -1 .. max_id -> values 0 .. max_id + 1
        int w = 1;
        while (w < 63 && (1LL << w) < card) w++;
        bits[c] = w;
        total += w;
        if (fixed < 64 && card > (1LL << fixed)) fits = false;
    }
    if (fits) {
        for (int c = 0; c < n_key_cols; c++) bits[c] = fixed;
        if (rank_independent) *rank_independent = 1;
        return MUSE_OK;
    }
    if (total > 64)
        return fail(MUSE_ERR_UNSUPPORTED, "the label cardinalities of the %d group-by keys need %d key bits (at most 64)", n_key_cols, total);
    if (rank_independent) *rank_independent = 0;
    return MUSE_OK;
}

static int setup_group_table(muse_batch *b, const RunArgs &a, KeyCols &kc, GroupTable &gt) {
    muse_group *g = b->g;
    int rc = key_bits(g, a.key_cols, a.n_key_cols, kc.bits, nullptr);
    if (rc) return rc;
    kc.ncols = a.n_key_cols;
    long double prod = 1;
    for (int c = 0; c < a.n_key_cols; c++) {
        const int col = a.key_cols[c];
        kc.col[c] = g->labels + (size_t)col * g->cap;
        const int64_t card = (int64_t)g->max_id[col] + 2;
        gt.radix[c] = card;
        prod *= (long double)card;
    }
    const int64_t S = g->size;
    const int64_t dense_limit = std::max<int64_t>(4 * S, 1 << 20);
    int64_t slots;
    if (prod <= (long double)dense_limit) {
        gt.dense = 1;
        slots = (int64_t)prod;
    } else {
        gt.dense = 0;
        slots = 1;
        while (slots < 2 * S) slots <<= 1;
    }
    if (slots > b->table_cap) {
        cudaFree(b->d_gmax); cudaFree(b->d_gidx); cudaFree(b->d_hkeys);
        b->d_gmax = b->d_hkeys = nullptr; b->d_gidx = nullptr; b->table_cap = 0;
        CU(cudaMalloc(&b->d_gmax, sizeof(unsigned long long) * (size_t)slots));
        CU(cudaMalloc(&b->d_hkeys, sizeof(unsigned long long) * (size_t)slots));
        CU(cudaMalloc(&b->d_gidx, sizeof(int32_t) * (size_t)slots));
        b->table_cap = slots;
    }
    gt.slots = slots;
    gt.gmax = b->d_gmax;
    gt.gidx = b->d_gidx;
    gt.hkeys = b->d_hkeys;
    cudaStream_t st = bstream(b);
    CU(cudaMemsetAsync(gt.gmax, 0, sizeof(unsigned long long) * (size_t)slots, st));
    CU(cudaMemsetAsync(gt.gidx, 0x7f, sizeof(int32_t) * (size_t)slots, st));
    if (!gt.dense) CU(cudaMemsetAsync(gt.hkeys, 0, sizeof(unsigned long long) * (size_t)slots, st));
    return MUSE_OK;
}

// fused screened runs launch the exact kernel over at most this many listed series without knowing
// the list length on the host; run_fused_overflow finishes a longer list after the fact
#define MUSE_EXACT_UB 32768
static int run_fused_overflow(muse_batch *b, int64_t n_exact, bool refill);

// Produces the selected records on the host (unsorted).  apply_filter == 0 keeps every
// representative (grouped multi-GPU partials).
static int run_select(muse_batch *b, const RunArgs &a, int apply_filter, int64_t limit, std::vector<Rec> &recs) {
    muse_group *g = b->g;
    const int64_t S = g->size;
    cudaStream_t st = bstream(b);
    recs.clear();
    if (S == 0) return MUSE_OK;
    const unsigned blocks = (unsigned)((S + 255) / 256);
    GroupTable gt;
    memset(&gt, 0, sizeof(gt));
    KeyCols kc;
    memset(&kc, 0, sizeof(kc));
    const bool grouped = a.n_key_cols > 0;
    if (grouped) {
        int rc = setup_group_table(b, a, kc, gt);
        if (rc) return rc;
        group_max_kernel<<<blocks, 256, 0, st>>>(gt, kc, b->d_score, S, b->d_slot);
        group_rep_kernel<<<blocks, 256, 0, st>>>(gt, b->d_score, S, b->d_slot);
        b->timing.n_launches += 2;
    }
    CU(cudaMemsetAsync(b->d_counters, 0, sizeof(unsigned long long) * 2, st));
    FilterArgs f{a.max_lag, a.threshold, a.sign_filter, apply_filter};
    Cand cand{b->d_ckey, b->d_cidx, b->d_clag, b->d_counters};
    emit_candidates_kernel<<<blocks, 256, 0, st>>>(gt, grouped ? b->d_slot : nullptr, b->d_score, b->d_lag, S, f, cand);
    b->timing.n_launches++;
    CU(cudaGetLastError());
    // one round trip: the count and the first CH records together (pinned mailbox)
    const size_t CH = std::min<size_t>((size_t)S, 32768);
    unsigned long long *h_n = reinterpret_cast<unsigned long long *>(b->h_pin);
    unsigned long long *h_key = reinterpret_cast<unsigned long long *>(b->h_pin + 64);
    int32_t *h_idx = reinterpret_cast<int32_t *>(b->h_pin + 64 + CH * 8);
    int32_t *h_lag = reinterpret_cast<int32_t *>(b->h_pin + 64 + CH * 12);
    CU(cudaMemcpyAsync(h_n, b->d_counters, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_key, b->d_ckey, sizeof(unsigned long long) * CH, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_idx, b->d_cidx, sizeof(int32_t) * CH, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_lag, b->d_clag, sizeof(int32_t) * CH, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsigned long long ncand = h_n[0];
    if (b->fused_run) {
        b->timing.n_rescored = (int64_t)h_n[2];
        b->timing.n_refined = (int64_t)h_n[3];
        if (b->fused_run == 1 && (int64_t)h_n[2] > std::min<int64_t>(S, MUSE_EXACT_UB)) {
            // rare: more exact candidates than the fixed launch covered -> finish them and select again
            const int64_t n_exact = (int64_t)h_n[2];
            int rc = run_fused_overflow(b, n_exact, false);
            if (rc) return rc;
            b->fused_run = 0;
            rc = run_select(b, a, apply_filter, limit, recs);
            b->timing.n_rescored = n_exact;
            return rc;
        }
    } else {
        b->timing.n_rescored = (int64_t)(h_n[2] + h_n[3]);     // screened runs: pilot + second round
    }
    if (ncand == 0) return MUSE_OK;
    auto better = [](const Rec &x, const Rec &y) { return x.key != y.key ? x.key > y.key : x.idx < y.idx; };
    if (ncand <= CH) {
        recs.resize((size_t)ncand);
        for (size_t i = 0; i < (size_t)ncand; i++) recs[i] = Rec{h_key[i], h_idx[i], h_lag[i]};
    } else {
        const unsigned long long *src_key = b->d_ckey;
        const int32_t *src_idx = b->d_cidx, *src_lag = b->d_clag;
        unsigned long long nsel = ncand;
        if (limit >= 0 && ncand > (unsigned long long)limit && ncand > 262144ull) {
            // device radix select of the `limit` best by (|score| desc, index asc)
            if (limit == 0) return MUSE_OK;
            CU(cudaMemsetAsync(b->d_sel, 0, sizeof(SelectState), st));
            unsigned long long want = (unsigned long long)limit;
            CU(cudaMemcpyAsync(&b->d_sel->want, &want, sizeof(want), cudaMemcpyHostToDevice, st));
            const unsigned sblocks = (unsigned)((ncand + 1023ull) / 1024ull);
            for (int r = 0; r < 6; r++) select_round_kernel<<<sblocks, 1024, 0, st>>>(b->d_sel, b->d_ckey, b->d_cidx, ncand, r);
            select_gather_kernel<<<sblocks, 1024, 0, st>>>(b->d_sel, b->d_ckey, b->d_cidx, ncand, b->d_clag, b->d_skey, b->d_sidx,
                                                           b->d_slag, b->d_counters + 1);
            b->timing.n_launches += 7;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(&nsel, b->d_counters + 1, sizeof(nsel), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            src_key = b->d_skey;
            src_idx = b->d_sidx;
            src_lag = b->d_slag;
        }
        std::vector<unsigned long long> hk((size_t)nsel);
        std::vector<int32_t> hi((size_t)nsel), hl((size_t)nsel);
        CU(cudaMemcpyAsync(hk.data(), src_key, sizeof(unsigned long long) * (size_t)nsel, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hi.data(), src_idx, sizeof(int32_t) * (size_t)nsel, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hl.data(), src_lag, sizeof(int32_t) * (size_t)nsel, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        recs.resize((size_t)nsel);
        for (size_t i = 0; i < (size_t)nsel; i++) recs[i] = Rec{hk[i], hi[i], hl[i]};
    }
    if (limit >= 0 && recs.size() > (size_t)limit) {
        std::nth_element(recs.begin(), recs.begin() + limit, recs.end(), better);
        recs.resize((size_t)limit);
    }
    std::sort(recs.begin(), recs.end(), better);
    return MUSE_OK;
}

static cudaError_t launch_screen(const muse_batch *b, const ScreenParams &p, cudaStream_t st) {
    if (b->log2m == 10) return launch_screen_warp(p, b->ctx->sm_count, st);
    // MUSE_WIDE13=1: the 512-thread, 64-register variant for n = 16384 (muse_screen_wide.cuh).  Measured SLOWER than the
    // 256-thread kernel (56.8 ms against 45.9 ms per 1.25 M x 10080): both issue at 39 % of peak, so its 24 % more
    // instructions decide; kept as the measured record of that design, off by default.
    if (b->log2m == 13 && b->twide_f && getenv("MUSE_WIDE13")) {
        ScreenParams pw = p;
        pw.twi = b->twide_f;
        return launch_screen_wide(pw, b->ctx->sm_count, st);
    }
    if (screen_is_big(b->log2m)) return launch_screen_big(b->log2m, p, b->ctx->sm_count, st);
    if (b->log2m == 6) return launch_screen_sub1(p, b->ctx->sm_count, st);
    if (b->log2m == 7) return launch_screen_sub2(p, b->ctx->sm_count, st);
    if (b->log2m == 8 && !block_small()) return launch_screen_sub3(p, b->ctx->sm_count, st);
    if (b->log2m == 9 && !block_small()) return launch_screen_sub4(p, b->ctx->sm_count, st);
    return launch_screen_block(b->log2m, p, b->ctx->sm_count, st);
}

static int ensure_lower(muse_batch *b) {
    const int64_t S = b->g->size;
    if (!b->d_L || b->d_L_cap < S) {
        cudaFree(b->d_L);
        cudaFree(b->d_W);
        b->d_L = nullptr;
        b->d_W = nullptr;
        b->d_L_cap = 0;
        CU(cudaMalloc(&b->d_L, sizeof(float) * (size_t)S));
        CU(cudaMalloc(&b->d_W, (size_t)S));
        b->d_L_cap = S;
    }
    return MUSE_OK;
}

// Per-row statistics of the store (RowStat): the ingest paths queue them behind the copy that brings the rows in
// (queue_row_stats, synth_rows_kernel); this is the safety net for rows that arrived any other way.
static int refresh_row_stats(muse_group *g) {
    if (g->stats_upto < g->size) {
        int rc = queue_row_stats(g, g->stats_upto, g->size - g->stats_upto);
        if (rc) return rc;
        g->stats_upto = g->size;
    }
    return MUSE_OK;
}

static ScreenParams screen_params(muse_batch *b) {
    ScreenParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.slab = b->g->slab;
    sp.ld = b->g->ld;
    sp.count = b->g->size;
    sp.N = (int)b->N;
    sp.twp = b->twp_f;
    sp.sw = b->sw_f;
    sp.a_mid = b->a_mid;
    sp.out_U = b->d_U;
    sp.sx = b->sx_f;
    sp.sb = b->sb_f;
    sp.twi = b->twi_f;
    sp.x_mid = b->x_mid;
    sp.row_stat = b->g->row_stat;
    sp.cut_bits = b->d_cut;
    sp.n_refined = reinterpret_cast<unsigned long long *>(b->d_cut + 2);
    sp.cut_hist = b->d_cut + 4;
    sp.top_n = 1;
    return sp;
}

// Arms the fused refinement of the warp kernel: running cut-off = cut0 (+inf: bounds only),
// counters and histogram cleared, lag window in the kernel's rotated cc index.
__global__ void init_cut_kernel(unsigned *state, float cut0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 + MUSE_CUT_WORDS) state[i] = (i == 0) ? __float_as_uint(cut0) : 0u;
}

// the same for the queries of a multi-query launch: blockIdx.y = query
__global__ void init_cut_batch_kernel(unsigned *const *__restrict__ states, float cut0) {
    unsigned *state = states[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 + MUSE_CUT_WORDS) state[i] = (i == 0) ? __float_as_uint(cut0) : 0u;
}

// skip_init: the cut-off state is armed by init_cut_batch_kernel for all the queries of a launch at once
static int arm_refinement(muse_batch *b, ScreenParams &sp, float cut0, int64_t max_lag, int64_t top_n, double threshold, bool skip_init = false) {
    if (!screen_is_fused(b->log2m)) return MUSE_OK;
    int rc = refresh_row_stats(b->g);
    if (rc) return rc;
    sp.row_stat = b->g->row_stat;
    if (!skip_init) {
        init_cut_kernel<<<(4 + MUSE_CUT_WORDS + 255) / 256, 256, 0, bstream(b)>>>(b->d_cut, cut0);
        CU(cudaGetLastError());
    }
    const int64_t n = b->n, pad = n - b->N;
    if (max_lag < 0) max_lag = -1;                          // nothing passes |lag| <= max_lag
    if (2 * max_lag >= n - 1) {                             // every lag is inside
        sp.win_lo = 0;
        sp.win_len = (int)n;
    } else {
        sp.win_lo = (int)(((pad - max_lag) % n + n) % n);
        sp.win_len = (int)(2 * max_lag);                    // -2 when max_lag < 0: no index qualifies
    }
    sp.top_n = (int)std::min<int64_t>(std::max<int64_t>(top_n, 1), 0x7fffffff);
    float thr = (float)threshold;
    if ((double)thr < threshold) thr = nextafterf(thr, INFINITY);    // round UP: a counted lower bound must really reach the threshold
    sp.thr = thr > 0.f ? thr : 0.f;
    return MUSE_OK;
}

// idx list of every series whose (refined) bound reaches the final cut-off, read from device memory
__global__ void survivors_cut_kernel(const float *__restrict__ U, int64_t S, const unsigned *__restrict__ cut_bits,
                                     int32_t *__restrict__ out, unsigned long long *n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) n[1] = *reinterpret_cast<const unsigned long long *>(cut_bits + 2);      // the refined count rides along (counters[3])
    const float lo = __uint_as_float(*cut_bits);
    const bool take = i < S && U[i] >= lo;
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    if (mask == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(n, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (take) out[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)i;
}

// the batched tails of muse_multi_run: blockIdx.y = query (MultiTail, muse_select.cuh)
__global__ void tail_reset_kernel(const MultiTail *__restrict__ tails, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq)
        for (int i = 0; i < 4; i++) tails[q].counters[i] = 0ull;
}
__global__ void survivors_cut_batch_kernel(const MultiTail *__restrict__ tails, int64_t S) {
    const MultiTail t = tails[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) t.counters[3] = *reinterpret_cast<const unsigned long long *>(t.cut + 2);
    const float lo = __uint_as_float(*t.cut);
    const bool take = i < S && t.U[i] >= lo;
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    if (mask == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(t.counters + 2, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (take) t.list[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)i;
}
// reference preparation of a launch's queries: screen_tables_kernel with blockIdx.y = query; the 4 floats of every query
// (std-zero flag, A[M/2], Xt[M/2]) land in ONE array for one copy to the host
struct TablesArgs {
    const cd *Xt;
    float4 *sw, *sx;
    const int32_t *std_zero;
    float *sb;
};
__global__ void screen_tables_batch_kernel(const TablesArgs *__restrict__ args, int M, const float2 *__restrict__ swtw, float *__restrict__ mids) {
    const TablesArgs a = args[blockIdx.y];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M / 2) return;
    float *mid = mids + 4 * blockIdx.y;
    if (k == 0) mid[0] = __int_as_float(*a.std_zero);
    const cd x = a.Xt[k], c = a.Xt[M - k];
    if (k == M / 2) {
        mid[1] = __double2float_ru(hypot(x.x, x.y) * 2.0);
        mid[2] = (float)x.x;
        mid[3] = (float)x.y;
        return;
    }
    const float wa = __double2float_ru(hypot(x.x, x.y) * (k == 0 ? 1.0 : 2.0));
    const float wc = __double2float_ru(hypot(c.x, c.y) * (k == 0 ? 1.0 : 2.0));
    const float2 w = swtw[k];
    a.sw[k] = make_float4(w.x, w.y, wa, wc);
    a.sx[k] = make_float4((float)x.x, (float)x.y, (float)c.x, (float)c.y);
    a.sb[k] = __double2float_ru(sqrt(2.0 * ((double)wa * (double)wa + (double)wc * (double)wc)) * (1.0 + 1e-15));
}

extern "C" int muse_batch_screen_bounds(muse_batch *b, int32_t refine, int64_t max_lag, float *upper, float *lower) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!upper || (refine && !lower)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_screen_bounds: NULL output");
    if (!b->screen_ok) return fail(MUSE_ERR_UNSUPPORTED, "no screening kernel for series length %lld", (long long)b->N);
    if (refine && !screen_is_fused(b->log2m)) return fail(MUSE_ERR_UNSUPPORTED, "the fused refinement needs an FFT length of 128 .. 16384");
    CU(cudaSetDevice(b->ctx->device));
    rc = ensure_scratch(b);
    if (rc) return rc;
    const int64_t S = b->g->size;
    if (S == 0) return MUSE_OK;
    cudaStream_t st = bstream(b);
    ScreenParams sp = screen_params(b);
    if (refine) {
        rc = ensure_lower(b);
        if (rc) return rc;
        sp.out_L = b->d_L;
    }
    // refine: cut-off 0 that never rises (top_n = INT_MAX): every series takes the second stage
    rc = arm_refinement(b, sp, refine ? 0.f : INFINITY, max_lag, 0x7fffffff, 0.0);
    if (rc) return rc;
    CU(launch_screen(b, sp, st));
    CU(cudaMemcpyAsync(upper, b->d_U, sizeof(float) * (size_t)S, cudaMemcpyDeviceToHost, st));
    if (refine) CU(cudaMemcpyAsync(lower, b->d_L, sizeof(float) * (size_t)S, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MUSE_OK;
}

// Screened scoring with the fused kernel (n = 2048, ungrouped): ONE pass over the slab gives
// every series an upper bound that is either the loose spectral one (below the running cut-off:
// cannot be in the result) or the tight fp32 one; the exact fp64 kernel then scores only the
// series whose bound reaches the final cut-off.  No host round trip before the final one: the
// cut-off and the list length stay on the device, the exact kernel is launched over
// MUSE_EXACT_UB series and skips what the list does not hold.
static int score_fused(muse_batch *b, const RunArgs &a) {
    const int64_t S = b->g->size;
    cudaStream_t st = bstream(b);
    ScreenParams sp = screen_params(b);
    const float thr_lo = a.threshold > 0 ? (float)a.threshold * 0.999999f : 0.f;   // never above the fp64 threshold
    int rc = MUSE_OK;
    if (b->prescreened) {      // muse_multi_run: this query's bounds and cut-off came out of the multi-query launch
        b->prescreened = 0;
    } else {
        rc = arm_refinement(b, sp, thr_lo, a.max_lag, a.top_n, a.threshold);
        if (rc) return rc;
        CU(launch_screen(b, sp, st));
        b->timing.n_launches += 2;
    }
    TIMING_EVENT(b, 1, st);
    if (!a.list_only) CU(cudaMemsetAsync(b->d_score, 0xff, sizeof(double) * (size_t)S, st));      // NaN = "cannot be in the result"
    const unsigned blocks = (unsigned)((S + 255) / 256);
    survivors_cut_kernel<<<blocks, 256, 0, st>>>(b->d_U, S, b->d_cut, b->d_list, b->d_counters + 2);
    b->timing.n_launches++;
    rc = score_exact_all(b, b->signed_run, b->d_list, std::min<int64_t>(S, MUSE_EXACT_UB), b->d_counters + 2);
    if (rc) return rc;
    return MUSE_OK;
}

// Grouped runs on the fused kernels: every series takes the fp32 second stage (bounds on the score
// itself, window ignored), a group's best LOWER bound prunes its members, and only the members whose
// upper bound reaches it -- the representative and whatever ties it within the fp32 slack -- are
// scored in fp64.  The group max / representative / filter / top-N then run on those exact scores
// exactly as in an all-exact run (every other member is provably below its group's representative).
static int score_fused_grouped(muse_batch *b, const RunArgs &a) {
    const int64_t S = b->g->size;
    cudaStream_t st = bstream(b);
    ScreenParams sp = screen_params(b);
    int rc = ensure_lower(b);
    if (rc) return rc;
    sp.out_L = b->d_L;
    sp.grouped = 1;
    // a member below the threshold cannot be the representative of a group that passes results.go:46-52
    const float thr_lo = a.threshold > 0 ? (float)a.threshold * 0.999999f : 0.f;
    rc = arm_refinement(b, sp, thr_lo, a.max_lag, 0x7fffffff, 0.0);      // a cut-off that never rises
    if (rc) return rc;
    GroupTable gt;
    memset(&gt, 0, sizeof(gt));
    KeyCols kc;
    memset(&kc, 0, sizeof(kc));
    rc = setup_group_table(b, a, kc, gt);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((S + 255) / 256);
    const bool running = screen_is_big(b->log2m);      // the kernel itself keeps each group's best lower bound
    if (running) {
        group_slots_kernel<<<blocks, 256, 0, st>>>(gt, kc, S, b->d_slot);
        sp.slot_of = b->d_slot;
        sp.group_L = gt.gmax;
        sp.out_W = b->d_W;
        b->timing.n_launches++;
    }
    CU(launch_screen(b, sp, st));
    b->timing.n_launches += 2;
    TIMING_EVENT(b, 1, st);
    CU(cudaMemsetAsync(b->d_score, 0xff, sizeof(double) * (size_t)S, st));      // NaN = "not its group's representative"
    if (!running) {
        group_lower_bound_kernel<<<blocks, 256, 0, st>>>(gt, kc, b->d_L, S, b->d_slot);
        b->timing.n_launches++;
    }
    const unsigned *cut_bits = nullptr;
    if (running && a.group_cut && a.top_n > 0 && a.top_n < 0x7fffffff && !getenv("MUSE_NO_GROUP_CUT")) {
        group_uncertain_kernel<<<blocks, 256, 0, st>>>(gt, b->d_U, b->d_W, S, b->d_slot);
        group_cut_count_kernel<<<(unsigned)((gt.slots + 255) / 256), 256, 0, st>>>(gt, a.threshold, b->d_cut + 4);
        group_cut_find_kernel<<<1, 32, 0, st>>>(b->d_cut + 4, (int)a.top_n, b->d_cut);
        b->timing.n_launches += 3;
        cut_bits = b->d_cut;
    }
    group_contenders_kernel<<<blocks, 256, 0, st>>>(gt, b->d_U, S, b->d_slot, thr_lo, b->d_list, b->d_counters + 2, cut_bits);
    b->timing.n_launches++;
    CU(cudaGetLastError());
    unsigned long long *h_n = reinterpret_cast<unsigned long long *>(b->h_pin);
    CU(cudaMemcpyAsync(h_n, b->d_counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));                                 // one small round trip: the list length sizes the launch
    const int64_t n_exact = (int64_t)h_n[0];
    rc = score_exact_all(b, b->signed_run, b->d_list, n_exact);
    if (rc) return rc;
    CU(cudaMemcpyAsync(b->d_counters + 3, b->d_cut + 2, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return MUSE_OK;
}

// The exact list was longer than the launch bound: score the rest now that its length is known.
// refill: the run skipped the NaN fill of the score array (list_only) and the caller is about to
// read it as a whole -- fill it and score the complete list again.
static int run_fused_overflow(muse_batch *b, int64_t n_exact, bool refill) {
    const int64_t S = b->g->size;
    if (refill) {
        CU(cudaMemsetAsync(b->d_score, 0xff, sizeof(double) * (size_t)S, bstream(b)));
        return score_exact_all(b, b->signed_run, b->d_list, n_exact);
    }
    if (n_exact <= std::min<int64_t>(S, MUSE_EXACT_UB)) return MUSE_OK;
    return score_exact_all(b, b->signed_run, b->d_list + MUSE_EXACT_UB, n_exact - MUSE_EXACT_UB);
}

static int run_scores(muse_batch *b, const RunArgs &a) {
    int rc = ensure_scratch(b);
    if (rc) return rc;
    memset(&b->timing, 0, sizeof(b->timing));
    b->fused_run = 0;
    b->timing_pending = 0;
    b->signed_run = a.signed_scores ? 1 : 0;
    cudaStream_t st = bstream(b);
    TIMING_EVENT(b, 0, st);
    CU(cudaMemsetAsync(b->d_counters, 0, sizeof(unsigned long long) * 4, st));
    // screening needs: a kernel for this FFT size, an ungrouped unsigned run, a sign filter that
    // unsigned scores can pass, and a store big enough to be worth two extra round trips
    // grouped runs are screened only by the fused kernels (the bound of a member says nothing about the
    // lag of its group's representative; the fused second stage bounds every member's score tightly)
    // signed scores (Muse.Run, muse.go:72-76): the bounds are bounds on |score|, which is what the filter's threshold and the
    // top-N rank by; the survivors are re-scored signed.  A sign filter on signed scores needs the sign, which no bound has.
    const bool can_screen = b->screen_ok && (a.n_key_cols == 0 || screen_is_fused(b->log2m)) &&
                            (a.signed_scores ? a.sign_filter == MUSE_SIGN_ANY : a.sign_filter != MUSE_SIGN_NEG);
    bool screen = can_screen && (a.mode == MUSE_MODE_SCREEN || (a.mode == MUSE_MODE_AUTO && b->g->size >= 16384));
    if (screen && a.n_key_cols > 0) {
        b->timing.mode = MUSE_MODE_SCREEN;
        b->fused_run = 2;          // statistics as a fused run; its exact list is already complete
        rc = score_fused_grouped(b, a);
        if (rc) return rc;
        TIMING_EVENT(b, 2, st);
        return MUSE_OK;
    }
    if (screen) {
        b->timing.mode = MUSE_MODE_SCREEN;
        b->fused_run = 1;
        rc = score_fused(b, a);
        if (rc) return rc;
        TIMING_EVENT(b, 2, st);
        return MUSE_OK;
    }
    b->timing.mode = MUSE_MODE_EXACT;
    rc = score_exact_all(b, a.signed_scores, nullptr, b->g->size);
    if (rc) return rc;
    TIMING_EVENT(b, 1, st);
    TIMING_EVENT(b, 2, st);
    return MUSE_OK;
}

// Elapsed times between the run's events.  The batches of a multi-query launch record none (TIMING_EVENT), so there is
// nothing to read for them; a failed query of an event must not stay behind as the runtime's last error.
static void read_timing_events(muse_batch *b) {
    float *out[4] = {&b->timing.total_ms, &b->timing.score_ms, &b->timing.rescore_ms, &b->timing.select_ms};
    const int from[4] = {0, 0, 1, 2}, to[4] = {3, 1, 2, 3};
    for (int i = 0; i < 4; i++)
        if (cudaEventElapsedTime(out[i], b->ev[from[i]], b->ev[to[i]]) != cudaSuccess) {
            (void)cudaGetLastError();
            *out[i] = 0.f;
        }
}

static int finish_timing(muse_batch *b) {
    if (b->use_aux) {      // a query of a multi-query launch: no events were recorded for this run
        CU(cudaStreamSynchronize(bstream(b)));
        return MUSE_OK;
    }
    cudaStream_t st = bstream(b);
    CU(cudaEventRecord(b->ev[3], st));
    CU(cudaEventSynchronize(b->ev[3]));
    read_timing_events(b);
    return MUSE_OK;
}

static int queue_topn_records(muse_batch *b, const RunArgs &a, muse_partial *d_out, int64_t capacity);

// Ungrouped runs with a device-side top-N, in two halves so that several batches can share ONE synchronisation
// (muse_multi_run): queue = scores, filter, rank, copy of the top_n records into the batch's pinned mailbox;
// finish (after the stream has been synchronised) = read the mailbox, or -- when the candidate list was too long for
// the device-side select, or the fused path's exact launch did not cover its list -- finish on the host path from
// the scores already on the device.
static bool device_topn_applies(const muse_batch *b, const RunArgs &a) {
    return a.n_key_cols == 0 && a.top_n > 0 && a.top_n <= 65536 && b->g->size > 0;
}

static int device_topn_queue(muse_batch *b, const RunArgs &a) {
    muse_partial *d_rec = reinterpret_cast<muse_partial *>(b->d_skey);      // free scratch in this path
    int rc = queue_topn_records(b, a, d_rec, a.top_n);
    if (rc) return rc;
    cudaStream_t st = bstream(b);
    CU(cudaMemcpyAsync(b->h_pin + 64, d_rec, sizeof(muse_partial) * (size_t)a.top_n, cudaMemcpyDeviceToHost, st));
    TIMING_EVENT(b, 3, st);
    return MUSE_OK;
}

static int device_topn_finish(muse_batch *b, const RunArgs &a, double *scores, int64_t *lags, int64_t *series_idx, int64_t *n_out) {
    const int64_t top_n = a.top_n;
    const muse_partial *h_rec = reinterpret_cast<const muse_partial *>(b->h_pin + 64);
    if (h_rec[0].flags != 2) {
        int64_t k = 0;
        for (; k < top_n && h_rec[k].flags == 0; k++) {
            scores[k] = h_rec[k].score;
            lags[k] = h_rec[k].lag;
            series_idx[k] = h_rec[k].series_idx;
        }
        *n_out = k;
        if (b->use_aux) return MUSE_OK;            // a query of a multi-query launch: no timing events were recorded
        muse_timing t;
        return muse_batch_last_timing(b, &t);      // settles timing_pending (events are complete)
    }
    b->timing_pending = 0;
    const unsigned long long *h_n = reinterpret_cast<const unsigned long long *>(b->h_pin);
    int rc;
    if (b->fused_run == 1) {
        rc = run_fused_overflow(b, (int64_t)h_n[2], true);
        if (rc) return rc;
        b->fused_run = 0;
    }
    std::vector<Rec> recs2;
    rc = run_select(b, a, 1, top_n, recs2);
    if (rc) return rc;
    for (size_t i = 0; i < recs2.size(); i++) {
        scores[i] = recs2[i].score();
        lags[i] = recs2[i].lag();
        series_idx[i] = b->g->global_offset + recs2[i].idx;
    }
    *n_out = (int64_t)recs2.size();
    return finish_timing(b);
}

extern "C" int muse_batch_run_ex(muse_batch *b, const int32_t *key_cols, int32_t n_key_cols, int64_t max_lag, int64_t top_n,
                                 double threshold, int32_t sign_filter, int32_t mode, int32_t signed_scores, double *scores,
                                 int64_t *lags, int64_t *series_idx, int64_t *n_out) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!n_out || (top_n > 0 && (!scores || !lags || !series_idx))) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run: NULL output");
    if (n_key_cols < 0 || (n_key_cols > 0 && !key_cols)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run: bad key columns");
    if (top_n < 0) top_n = 0;
    CU(cudaSetDevice(b->ctx->device));
    RunArgs a{key_cols, n_key_cols, max_lag, top_n, threshold, sign_filter, mode, signed_scores};
    a.group_cut = 1;
    *n_out = 0;
    if (device_topn_applies(b, a)) {
        // ungrouped: filter and top-N on the device, ONE copy of top_n records to the host
        rc = ensure_scratch(b);
        if (rc) return rc;
        if (top_n * 4 <= b->scratch_cap) {
            rc = device_topn_queue(b, a);
            if (rc) return rc;
            CU(cudaStreamSynchronize(bstream(b)));
            return device_topn_finish(b, a, scores, lags, series_idx, n_out);
        }
    }
    rc = run_scores(b, a);
    if (rc) return rc;
    std::vector<Rec> recs;
    rc = run_select(b, a, 1, top_n, recs);
    if (rc) return rc;
    for (size_t i = 0; i < recs.size(); i++) {
        scores[i] = recs[i].score();
        lags[i] = recs[i].lag();
        series_idx[i] = b->g->global_offset + recs[i].idx;
    }
    *n_out = (int64_t)recs.size();
    return finish_timing(b);
}

extern "C" int muse_batch_run(muse_batch *b, const int32_t *key_cols, int32_t n_key_cols, int64_t max_lag, int64_t top_n,
                              double threshold, int32_t sign_filter, double *scores, int64_t *lags, int64_t *series_idx,
                              int64_t *n_out) {
    return muse_batch_run_ex(b, key_cols, n_key_cols, max_lag, top_n, threshold, sign_filter, MUSE_MODE_AUTO, 0, scores, lags,
                             series_idx, n_out);
}

// ------------------------------------------------------------------------------------
// multi-GPU partials
// ------------------------------------------------------------------------------------
extern "C" int64_t muse_batch_partial_capacity(muse_batch *b, const int32_t *key_cols, int32_t n_key_cols, int64_t top_n) {
    (void)key_cols;
    if (!b) return 0;
    if (n_key_cols <= 0) return std::min<int64_t>(std::max<int64_t>(top_n, 0), b->g->size);
    return b->g->size;   // at most one representative per series
}

// Ungrouped shard partials without a host round trip: scores, filter and the shard's top_n stay on
// the device; `d_out` (DEVICE memory, `capacity` >= top_n records) is complete when the work queued
// on the context's stream has run.  Nothing is synchronised here, so a collective enqueued on the
// same stream (ncclAllGather of the records) follows directly.
// Queues (no synchronisation): scores, filter, device-side top_n of an UNGROUPED run as muse_partial
// records in device memory, the counters into the pinned mailbox, and the end-of-run event.
static int queue_topn_records(muse_batch *b, const RunArgs &a_in, muse_partial *d_out, int64_t capacity) {
    static_assert(sizeof(PartialRec) == sizeof(muse_partial), "PartialRec mirrors muse_partial");
    RunArgs a = a_in;
    a.list_only = 1;
    const int64_t top_n = a.top_n;
    int rc = run_scores(b, a);
    if (rc) return rc;
    const int64_t S = b->g->size;
    cudaStream_t st = bstream(b);
    // counters[0..1] (candidates, selected) are still zero: run_scores cleared all four and ungrouped scoring writes only [2..3]
    if (S > 0) {
        FilterArgs f{a.max_lag, a.threshold, a.sign_filter, 1};
        Cand cand{b->d_ckey, b->d_cidx, b->d_clag, b->d_counters};
        GroupTable gt;
        memset(&gt, 0, sizeof(gt));
        if (b->fused_run == 1) {
            // only the listed series have scores: filter those (the list's length stays on the device)
            const int64_t lim = std::min<int64_t>(S, MUSE_EXACT_UB);
            emit_listed_kernel<<<(unsigned)((lim + 255) / 256), 256, 0, st>>>(b->d_list, b->d_counters + 2, lim, b->d_score, b->d_lag, f, cand);
        } else {
            emit_candidates_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(gt, nullptr, b->d_score, b->d_lag, S, f, cand);
        }
        b->timing.n_launches++;
    }
    // the grid covers MUSE_PARTIAL_RANK_CAP candidates; blocks past the real count (device side) leave at once
    const long long exact_bound = b->fused_run == 1 ? (long long)std::min<int64_t>(S, MUSE_EXACT_UB) : -1;
    const unsigned pblocks = (unsigned)std::min<int64_t>((std::max<int64_t>(S, 1) + 7) / 8, (int64_t)b->ctx->sm_count * 8);   // 8 warps per block, one warp per candidate
    partial_topn_kernel<<<pblocks, 256, 0, st>>>(b->d_ckey, b->d_cidx, b->d_clag, b->d_counters, (long long)top_n,
                                                 (long long)b->g->global_offset, exact_bound, reinterpret_cast<PartialRec *>(d_out),
                                                 (long long)capacity);
    b->timing.n_launches++;
    CU(cudaGetLastError());
    // statistics and the end-of-run event are picked up by muse_batch_last_timing
    CU(cudaMemcpyAsync(b->h_pin, b->d_counters, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
    TIMING_EVENT(b, 3, st);
    b->timing_pending = b->use_aux ? 0 : 1;
    return MUSE_OK;
}

extern "C" int muse_batch_run_partial_device(muse_batch *b, int64_t max_lag, int64_t top_n, double threshold, int32_t sign_filter,
                                             int32_t mode, muse_partial *d_out, int64_t capacity) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!d_out || capacity < 1) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_partial_device: no output records");
    if (top_n < 0) top_n = 0;
    if (top_n > capacity) return fail(MUSE_ERR_INVALID_ARG, "partial capacity %lld < top_n %lld", (long long)capacity, (long long)top_n);
    CU(cudaSetDevice(b->ctx->device));
    RunArgs a{nullptr, 0, max_lag, top_n, threshold, sign_filter, mode, 0};
    return queue_topn_records(b, a, d_out, capacity);
}

static int host_keys(muse_batch *b, const RunArgs &a, const std::vector<Rec> &recs, std::vector<uint64_t> &keys) {
    // canonical group key of each record's series (labels fetched per record)
    keys.resize(recs.size());
    muse_group *g = b->g;
    int bits[MUSE_MAX_KEY_COLS], rank_independent = 0;
    int rc = key_bits(g, a.key_cols, a.n_key_cols, bits, &rank_independent);
    if (rc) return rc;
    if (!rank_independent)
        return fail(MUSE_ERR_UNSUPPORTED, "shard partials need group keys of at most %d bits per column (64 / %d keys)", 64 / a.n_key_cols, a.n_key_cols);
    cudaStream_t st = bstream(b);
    std::vector<int32_t> ids(recs.size() * (size_t)a.n_key_cols);
    if (recs.size() > 4096) {
        std::vector<int32_t> col((size_t)g->size);
        for (int c = 0; c < a.n_key_cols; c++) {
            CU(cudaMemcpyAsync(col.data(), g->labels + (size_t)a.key_cols[c] * g->cap, sizeof(int32_t) * (size_t)g->size, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (size_t i = 0; i < recs.size(); i++) ids[i * a.n_key_cols + c] = col[(size_t)recs[i].idx];
        }
    } else {
        for (size_t i = 0; i < recs.size(); i++)
            for (int c = 0; c < a.n_key_cols; c++)
                CU(cudaMemcpyAsync(&ids[i * a.n_key_cols + c], g->labels + (size_t)a.key_cols[c] * g->cap + recs[i].idx, sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    for (size_t i = 0; i < recs.size(); i++) {
        uint64_t key = 0;
        for (int c = 0; c < a.n_key_cols; c++) {
            const uint64_t v = (uint64_t)(ids[i * a.n_key_cols + c] + 1);
            key = bits[c] >= 64 ? v : ((key << bits[c]) | v);
        }
        keys[i] = key;
    }
    return MUSE_OK;
}

extern "C" int muse_batch_run_partial(muse_batch *b, const int32_t *key_cols, int32_t n_key_cols, int64_t max_lag, int64_t top_n,
                                      double threshold, int32_t sign_filter, int32_t mode, muse_partial *out, int64_t capacity,
                                      int64_t *n_out) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!n_out || (!out && capacity > 0)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_partial: NULL output");
    if (n_key_cols < 0 || (n_key_cols > 0 && !key_cols)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_partial: bad key columns");
    if (top_n < 0) top_n = 0;
    CU(cudaSetDevice(b->ctx->device));
    RunArgs a{key_cols, n_key_cols, max_lag, top_n, threshold, sign_filter, mode, 0};
    *n_out = 0;
    rc = run_scores(b, a);
    if (rc) return rc;
    std::vector<Rec> recs;
    const bool grouped = n_key_cols > 0;
    // grouped: every representative, unfiltered (the filter needs the GLOBAL group max);
    // ungrouped: the shard's own filtered top_n is enough
    rc = run_select(b, a, grouped ? 0 : 1, grouped ? -1 : top_n, recs);
    if (rc) return rc;
    if ((int64_t)recs.size() > capacity) return fail(MUSE_ERR_INVALID_ARG, "partial capacity %lld < %zu records", (long long)capacity, recs.size());
    std::vector<uint64_t> keys;
    if (grouped) {
        rc = host_keys(b, a, recs, keys);
        if (rc) return rc;
    }
    for (size_t i = 0; i < recs.size(); i++) {
        out[i].series_idx = b->g->global_offset + recs[i].idx;
        out[i].group_key = grouped ? keys[i] : (uint64_t)out[i].series_idx;
        out[i].score = recs[i].score();
        out[i].lag = recs[i].lag();
        out[i].flags = 0;
    }
    *n_out = (int64_t)recs.size();
    return finish_timing(b);
}

extern "C" int muse_merge_partials(const muse_partial *parts, int64_t n_parts, int64_t max_lag, int64_t top_n, double threshold,
                                   int32_t sign_filter, double *scores, int64_t *lags, int64_t *series_idx, int64_t *n_out) {
    if (!n_out || (n_parts > 0 && !parts)) return fail(MUSE_ERR_INVALID_ARG, "muse_merge_partials: NULL argument");
    if (top_n < 0) top_n = 0;
    if (top_n > 0 && (!scores || !lags || !series_idx)) return fail(MUSE_ERR_INVALID_ARG, "muse_merge_partials: NULL output");
    // group max across shards first (muse_batch.go:87-89), ties -> lowest global series index
    std::unordered_map<uint64_t, muse_partial> best;
    best.reserve((size_t)n_parts * 2 + 1);
    for (int64_t i = 0; i < n_parts; i++) {
        const muse_partial &p = parts[i];
        if ((p.flags & 1) || p.score != p.score) continue;
        auto it = best.find(p.group_key);
        if (it == best.end()) best.emplace(p.group_key, p);
        else {
            muse_partial &q = it->second;
            if (fabs(p.score) > fabs(q.score) || (fabs(p.score) == fabs(q.score) && p.series_idx < q.series_idx)) q = p;
        }
    }
    std::vector<muse_partial> v;
    v.reserve(best.size());
    for (auto &kv : best) {
        const muse_partial &p = kv.second;
        const int64_t al = p.lag < 0 ? -(int64_t)p.lag : (int64_t)p.lag;   // results.go:46-52
        const bool ok = al <= max_lag && fabs(p.score) >= threshold &&
                        (sign_filter == 0 || (p.score > 0 && sign_filter == 1) || (p.score < 0 && sign_filter == -1));
        if (ok) v.push_back(p);
    }
    auto better = [](const muse_partial &x, const muse_partial &y) {
        return fabs(x.score) != fabs(y.score) ? fabs(x.score) > fabs(y.score) : x.series_idx < y.series_idx;
    };
    if ((int64_t)v.size() > top_n) {
        std::nth_element(v.begin(), v.begin() + top_n, v.end(), better);
        v.resize((size_t)top_n);
    }
    std::sort(v.begin(), v.end(), better);
    for (size_t i = 0; i < v.size(); i++) {
        scores[i] = v[i].score;
        lags[i] = v[i].lag;
        series_idx[i] = v[i].series_idx;
    }
    *n_out = (int64_t)v.size();
    return MUSE_OK;
}

// ------------------------------------------------------------------------------------
// many reference queries against one resident store (BASELINE.json configs[4])
// ------------------------------------------------------------------------------------
// Round-1 form: one NewBatch + Run per reference on the resident slab, inside the library -- the store, its
// row statistics and the run scratch (context pool) are shared, results of query q land in row q of the
// outputs.  Every query still streams the slab once (HBM-bound, 2.3 ms per 1 M x 1440 series); the multi-query
// bound pass that reads the slab once for ALL queries is DESIGN.md section 8's next step.
// Up to ScreenMultiCfg::QC queries in ONE pass over the slab: every batch gets its cut-off state armed, the
// multi-query kernel fills each batch's bounds (d_U) and cut-off, and each batch is marked `prescreened` so that
// its next fused run starts at the tail (survivors -> exact fp64 kernel -> filter -> top-N).
static int tc_bounds_queue(muse_ctx *ctx, muse_group *g, muse_batch **bs, int nq);

static int screen_multi(muse_ctx *ctx, muse_batch **bs, int nq, int64_t max_lag, int64_t top_n, double threshold, bool tensor_cores) {
    cudaStream_t st = ctx->stream;
    std::vector<MultiQuery> hq((size_t)nq);
    ScreenParams sp0;
    const float thr_lo = threshold > 0 ? (float)threshold * 0.999999f : 0.f;
    for (int q = 0; q < nq; q++) {
        muse_batch *b = bs[q];
        int rc = ensure_scratch(b);
        if (rc) return rc;
        ScreenParams sp = screen_params(b);
        rc = arm_refinement(b, sp, thr_lo, max_lag, top_n, threshold);
        if (rc) return rc;
        if (q == 0) sp0 = sp;
        MultiQuery &m = hq[(size_t)q];
        memset(&m, 0, sizeof(m));
        m.sw = b->sw_f;
        m.sx = b->sx_f;
        m.x_mid = b->x_mid;
        m.a_mid = b->a_mid;
        m.cut = b->d_cut;
        m.out_U = b->d_U;
    }
    // the query table lives with the context (stream-ordered reuse: copy and launch go down the same stream; a copy
    // from pageable memory has left the host buffer when the call returns), so nothing here waits for the kernel
    if (!ctx->d_multi_q) CU(cudaMalloc(&ctx->d_multi_q, sizeof(MultiQuery) * (size_t)MUSE_MULTI_GROUP));
    MultiQuery *d_q = static_cast<MultiQuery *>(ctx->d_multi_q);
    CU(cudaMemcpyAsync(d_q, hq.data(), sizeof(MultiQuery) * (size_t)nq, cudaMemcpyHostToDevice, st));
    if (tensor_cores) {
        // bounds of all queries as one bf16 contraction on the tensor cores, then ONE pass over the store for the second
        // stages of all of them
        int rc = tc_bounds_queue(ctx, bs[0]->g, bs, nq);
        if (rc) return rc;
        if (!ctx->d_multi_next) CU(cudaMalloc(&ctx->d_multi_next, sizeof(unsigned)));
        CU(launch_refine_multi(sp0, d_q, nq, ctx->sm_count, ctx->d_multi_next, st));
    } else {
        CU(launch_screen_multi(sp0, d_q, nq, ctx->sm_count, st));
    }
    for (int q = 0; q < nq; q++) bs[q]->prescreened = 1;
    return MUSE_OK;
}

// The bounds of up to TcCfg::TN queries (batches bs[0 .. nq), FFT length 2048, tables built) against the whole store, as
// one bf16 contraction on the tensor cores: magnitudes of every series -> tiles, the queries' weights -> tiles, GEMM with
// the bound's epilogue into every batch's d_U.  Everything is queued on the context's stream.
static int tc_bounds_queue(muse_ctx *ctx, muse_group *g, muse_batch **bs, int nq) {
    const int64_t S = g->size;
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < 5; i++)
        if (!ctx->mt_ev[i]) CU(cudaEventCreate(&ctx->mt_ev[i]));
    int rc = refresh_row_stats(g);
    if (rc) return rc;
    const size_t a_bytes = TcCfg::a_bytes(S);
    if (ctx->tc_a_bytes < a_bytes) {
        cudaFree(ctx->tc_a);
        ctx->tc_a = nullptr;
        ctx->tc_a_bytes = 0;
        CU(cudaMalloc(&ctx->tc_a, a_bytes));
        ctx->tc_a_bytes = a_bytes;
    }
    if (ctx->tc_mid_cap < S) {
        cudaFree(ctx->tc_mid);
        ctx->tc_mid = nullptr;
        ctx->tc_mid_cap = 0;
        CU(cudaMalloc(&ctx->tc_mid, sizeof(float) * (size_t)S));
        ctx->tc_mid_cap = S;
    }
    if (!ctx->tc_b) CU(cudaMalloc(&ctx->tc_b, TcCfg::b_bytes()));
    if (!ctx->tc_amid) CU(cudaMalloc(&ctx->tc_amid, sizeof(float) * TcCfg::TN));
    if (!ctx->tc_ptrs) CU(cudaMalloc(&ctx->tc_ptrs, sizeof(void *) * 2 * TcCfg::TN));
    void *h_ptrs[2 * TcCfg::TN];
    float h_amid[TcCfg::TN];
    memset(h_ptrs, 0, sizeof(h_ptrs));
    memset(h_amid, 0, sizeof(h_amid));
    for (int q = 0; q < nq; q++) {
        rc = ensure_scratch(bs[q]);
        if (rc) return rc;
        h_ptrs[q] = bs[q]->sw_f;
        h_ptrs[TcCfg::TN + q] = bs[q]->d_U;
        h_amid[q] = bs[q]->a_mid;
    }
    // pageable sources: the copies have left the host buffers when the calls return
    CU(cudaMemcpyAsync(ctx->tc_ptrs, h_ptrs, sizeof(h_ptrs), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->tc_amid, h_amid, sizeof(h_amid), cudaMemcpyHostToDevice, st));
    ScreenParams sp = screen_params(bs[0]);
    CU(cudaEventRecord(ctx->mt_ev[0], st));
    CU(launch_mag_tiles(sp, ctx->tc_a, ctx->tc_mid, ctx->sm_count, st));
    CU(launch_weight_tiles(reinterpret_cast<const float4 *const *>(ctx->tc_ptrs), nq, ctx->tc_b, st));
    CU(cudaEventRecord(ctx->mt_ev[1], st));
    TcBoundsParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.a_tiles = ctx->tc_a;
    tp.b_tiles = ctx->tc_b;
    tp.mid = ctx->tc_mid;
    tp.amid = ctx->tc_amid;
    tp.row_stat = g->row_stat;
    tp.out_U = reinterpret_cast<float *const *>(ctx->tc_ptrs + TcCfg::TN);
    tp.S = S;
    tp.nq = nq;
    CU(launch_bounds_tc(tp, st));
    CU(cudaEventRecord(ctx->mt_ev[2], st));
    return MUSE_OK;
}

static bool tc_shape_ok(const muse_group *g, int64_t ref_len) {
    return ref_len == g->N && next_pow2(ref_len) == 2048;
}

// Layout of the per-launch tables of the tensor-core multi-query path (the same in device memory and in the pinned host
// mirror): parameter blocks of the batched kernels, then the results that come back.
struct MultiTables {
    static constexpr int Q = MUSE_MULTI_GROUP;
    static size_t off_ep() { return 0; }                                                  // ExactParams[Q]: reference spectra
    static size_t off_ta() { return off_ep() + sizeof(ExactParams) * Q; }                 // TablesArgs[Q]
    static size_t off_cut() { return off_ta() + sizeof(TablesArgs) * Q; }                 // unsigned*[Q]: cut-off states
    static size_t off_tail() { return off_cut() + sizeof(void *) * Q; }                   // MultiTail[Q]
    static size_t off_ep2() { return off_tail() + sizeof(MultiTail) * Q; }                // ExactParams[Q]: fp64 re-scoring
    static size_t off_mids() { return off_ep2() + sizeof(ExactParams) * Q; }              // float[Q][4]
    static size_t off_cnt() { return off_mids() + sizeof(float) * 4 * Q; }                // u64[Q][4]
    static size_t off_recs() { return (off_cnt() + sizeof(unsigned long long) * 4 * Q + 255) / 256 * 256; }      // PartialRec[Q][top_n]
    static size_t bytes(int64_t top_n) { return off_recs() + sizeof(PartialRec) * (size_t)Q * (size_t)std::max<int64_t>(top_n, 1); }
};

// One launch group (<= 256 references) of muse_multi_run on the tensor-core path, with ONE launch per stage for all its
// queries: reference spectra and tables (batched fp64 kernel in MODE_REF, blockIdx.y = query), cut-off states, bounds
// (bf16 contraction), second stages (one pass over the store), survivors, fp64 re-scoring, filter, top-N -- and two
// round trips to the host: the std-zero flags after the preparation, the list lengths before the re-scoring.  A query
// whose lists do not fit the device-side select finishes alone on the single-query path.
static int multi_run_tc_group(muse_ctx *ctx, muse_group *g, const double *refs, int nq, int64_t ref_len, int64_t max_lag, int64_t top_n,
                              double threshold, int32_t sign_filter, double *scores, int64_t *lags, int64_t *series_idx, int64_t *n_out) {
    using MT = MultiTables;
    const int64_t S = g->size;
    cudaStream_t st = ctx->stream;
    if (!ctx->d_mt || ctx->mt_top_n < top_n) {
        cudaFree(ctx->d_mt);
        if (ctx->h_mt) cudaFreeHost(ctx->h_mt);
        ctx->d_mt = ctx->h_mt = nullptr;
        ctx->mt_top_n = 0;
        CU(cudaMalloc(&ctx->d_mt, MT::bytes(top_n)));
        CU(cudaHostAlloc((void **)&ctx->h_mt, MT::bytes(top_n), cudaHostAllocDefault));
        ctx->mt_top_n = top_n;
    }
    unsigned char *d = ctx->d_mt, *h = ctx->h_mt;
    if (ctx->d_multi_ld != g->ld || ctx->d_multi_n != g->N) {
        cudaFree(ctx->d_multi_refs);
        ctx->d_multi_refs = nullptr;
        ctx->d_multi_ld = ctx->d_multi_n = 0;
        CU(cudaMalloc(&ctx->d_multi_refs, sizeof(double) * (size_t)MUSE_MULTI_GROUP * (size_t)g->ld));
        CU(cudaMemsetAsync(ctx->d_multi_refs, 0, sizeof(double) * (size_t)MUSE_MULTI_GROUP * (size_t)g->ld, st));
        ctx->d_multi_ld = g->ld;
        ctx->d_multi_n = g->N;
    }
    CU(cudaMemcpy2DAsync(ctx->d_multi_refs, sizeof(double) * (size_t)g->ld, refs, sizeof(double) * (size_t)ref_len,
                         sizeof(double) * (size_t)ref_len, (size_t)nq, cudaMemcpyHostToDevice, st));
    std::vector<muse_batch *> bs((size_t)nq, nullptr);
    int rc = MUSE_OK;
    auto cleanup = [&](int code) {
        cudaStreamSynchronize(st);
        for (muse_batch *b : bs)
            if (b) muse_batch_destroy(b);
        return code;
    };
#define CUM(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess) {                                                                                         \
            (void)cudaGetLastError();                                                                                    \
            return cleanup(fail(e_ == cudaErrorMemoryAllocation ? MUSE_ERR_OUT_OF_MEMORY : MUSE_ERR_CUDA, "%s failed: %s (%s:%d)", \
                                #call, cudaGetErrorString(e_), __FILE__, __LINE__));                                      \
        }                                                                                                                \
    } while (0)

    // ---- reference preparation: one launch per stage ----
    ExactParams *h_ep = reinterpret_cast<ExactParams *>(h + MT::off_ep());
    TablesArgs *h_ta = reinterpret_cast<TablesArgs *>(h + MT::off_ta());
    unsigned **h_cut = reinterpret_cast<unsigned **>(h + MT::off_cut());
    for (int i = 0; i < nq; i++) {
        rc = batch_alloc(ctx, g, ref_len, &bs[(size_t)i]);
        if (rc == MUSE_OK) rc = ensure_scratch(bs[(size_t)i]);
        if (rc) return cleanup(rc);
        muse_batch *b = bs[(size_t)i];
        ExactParams &p = h_ep[i];
        memset(&p, 0, sizeof(p));
        p.slab = ctx->d_multi_refs + (size_t)i * (size_t)g->ld;
        p.ld = g->ld;
        p.count = 1;
        p.N = (int)ref_len;
        p.twM = b->twM;
        p.twn = b->twn;
        p.out_X = b->Xt;
        p.out_flag = b->d_flag;
        h_ta[i] = TablesArgs{b->Xt, b->sw_f, b->sx_f, b->d_flag, b->sb_f};
        h_cut[i] = b->d_cut;
    }
    const int64_t M = bs[0]->n / 2;
    const float thr_lo = threshold > 0 ? (float)threshold * 0.999999f : 0.f;
    CUM(cudaMemcpyAsync(d, h, MT::off_tail(), cudaMemcpyHostToDevice, st));
    CUM(launch_exact_batch(MODE_REF, reinterpret_cast<const ExactParams *>(d + MT::off_ep()), nq, 1, st));
    screen_tables_batch_kernel<<<dim3((unsigned)((M / 2 + 1 + 255) / 256), (unsigned)nq), 256, 0, st>>>(
        reinterpret_cast<const TablesArgs *>(d + MT::off_ta()), (int)M, bs[0]->swtw, reinterpret_cast<float *>(d + MT::off_mids()));
    init_cut_batch_kernel<<<dim3((4 + MUSE_CUT_WORDS + 255) / 256, (unsigned)nq), 256, 0, st>>>(
        reinterpret_cast<unsigned *const *>(d + MT::off_cut()), thr_lo);
    CUM(cudaGetLastError());
    CUM(cudaMemcpyAsync(h + MT::off_mids(), d + MT::off_mids(), sizeof(float) * 4 * (size_t)nq, cudaMemcpyDeviceToHost, st));
    CUM(cudaStreamSynchronize(st));
    const float *h_mids = reinterpret_cast<const float *>(h + MT::off_mids());
    std::vector<int> live;
    for (int i = 0; i < nq; i++) {
        int32_t flag;
        memcpy(&flag, &h_mids[4 * i], sizeof(flag));
        if (flag) {      // muse_batch.go:38-41: this query has no Batch; the others do
            n_out[i] = -1;
            continue;
        }
        muse_batch *b = bs[(size_t)i];
        b->a_mid = h_mids[4 * i + 1];
        b->x_mid = cf{h_mids[4 * i + 2], h_mids[4 * i + 3]};
        b->screen_ok = 1;
        n_out[i] = 0;
        live.push_back(i);
    }
    const int nl = (int)live.size();
    if (nl == 0) return cleanup(MUSE_OK);

    // ---- bounds (tensor cores) and second stages (one pass over the store) ----
    std::vector<muse_batch *> lb((size_t)nl);
    std::vector<MultiQuery> hq((size_t)nl);
    ScreenParams sp0;
    for (int k = 0; k < nl; k++) {
        muse_batch *b = lb[(size_t)k] = bs[(size_t)live[(size_t)k]];
        ScreenParams sp = screen_params(b);
        rc = arm_refinement(b, sp, thr_lo, max_lag, top_n, threshold, true);
        if (rc) return cleanup(rc);
        if (k == 0) sp0 = sp;
        MultiQuery &m = hq[(size_t)k];
        memset(&m, 0, sizeof(m));
        m.sw = b->sw_f;
        m.sx = b->sx_f;
        m.x_mid = b->x_mid;
        m.a_mid = b->a_mid;
        m.cut = b->d_cut;
        m.out_U = b->d_U;
    }
    if (!ctx->d_multi_q) CUM(cudaMalloc(&ctx->d_multi_q, sizeof(MultiQuery) * (size_t)MUSE_MULTI_GROUP));
    if (!ctx->d_multi_next) CUM(cudaMalloc(&ctx->d_multi_next, sizeof(unsigned)));
    MultiQuery *d_q = static_cast<MultiQuery *>(ctx->d_multi_q);
    CUM(cudaMemcpyAsync(d_q, hq.data(), sizeof(MultiQuery) * (size_t)nl, cudaMemcpyHostToDevice, st));
    rc = tc_bounds_queue(ctx, g, lb.data(), nl);
    if (rc) return cleanup(rc);
    CUM(launch_refine_multi(sp0, d_q, nl, ctx->sm_count, ctx->d_multi_next, st));
    CUM(cudaEventRecord(ctx->mt_ev[3], st));

    // ---- tails: survivors -> fp64 re-scoring -> filter -> top-N, one launch each ----
    MultiTail *h_tail = reinterpret_cast<MultiTail *>(h + MT::off_tail());
    ExactParams *h_ep2 = reinterpret_cast<ExactParams *>(h + MT::off_ep2());
    unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(d + MT::off_cnt());
    PartialRec *d_recs = reinterpret_cast<PartialRec *>(d + MT::off_recs());
    const int64_t lim = std::min<int64_t>(S, MUSE_EXACT_UB);
    for (int k = 0; k < nl; k++) {
        muse_batch *b = lb[(size_t)k];
        h_tail[k] = MultiTail{b->d_U, b->d_cut, b->d_list, d_cnt + 4 * k, b->d_score, b->d_lag, b->d_ckey, b->d_cidx, b->d_clag,
                              d_recs + (size_t)k * (size_t)top_n};
        ExactParams &p = h_ep2[k];
        memset(&p, 0, sizeof(p));
        p.slab = g->slab;
        p.ld = g->ld;
        p.count = lim;
        p.count_ptr = d_cnt + 4 * k + 2;
        p.idx = b->d_list;
        p.N = (int)ref_len;
        p.Xt = b->Xt;
        p.twM = b->twM;
        p.twn = b->twn;
        p.out_score = b->d_score;
        p.out_lag = b->d_lag;
    }
    CUM(cudaMemcpyAsync(d + MT::off_tail(), h + MT::off_tail(), MT::off_mids() - MT::off_tail(), cudaMemcpyHostToDevice, st));
    const MultiTail *d_tail = reinterpret_cast<const MultiTail *>(d + MT::off_tail());
    tail_reset_kernel<<<(nl + 255) / 256, 256, 0, st>>>(d_tail, nl);
    survivors_cut_batch_kernel<<<dim3((unsigned)((S + 255) / 256), (unsigned)nl), 256, 0, st>>>(d_tail, S);
    CUM(cudaGetLastError());
    unsigned long long *h_cnt = reinterpret_cast<unsigned long long *>(h + MT::off_cnt());
    CUM(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long) * 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
    CUM(cudaStreamSynchronize(st));      // the list lengths size the next launches
    int64_t max_len = 0;
    for (int k = 0; k < nl; k++) max_len = std::max<int64_t>(max_len, std::min<int64_t>((int64_t)h_cnt[4 * k + 2], lim));
    if (max_len > 0) {
        CUM(launch_exact_batch(MODE_SCORE, reinterpret_cast<const ExactParams *>(d + MT::off_ep2()), nl, max_len, st));
        FilterArgs f{max_lag, threshold, sign_filter, 1};
        emit_listed_batch_kernel<<<dim3((unsigned)((max_len + 255) / 256), (unsigned)nl), 256, 0, st>>>(d_tail, (long long)lim, f);
    }
    partial_topn_batch_kernel<<<dim3(16, (unsigned)nl), 256, 0, st>>>(d_tail, (long long)top_n, (long long)g->global_offset, (long long)lim);
    CUM(cudaGetLastError());
    CUM(cudaMemcpyAsync(h + MT::off_recs(), d_recs, sizeof(PartialRec) * (size_t)nl * (size_t)top_n, cudaMemcpyDeviceToHost, st));
    CUM(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long) * 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
    CUM(cudaEventRecord(ctx->mt_ev[4], st));
    CUM(cudaStreamSynchronize(st));
    const muse_partial *h_recs = reinterpret_cast<const muse_partial *>(h + MT::off_recs());
    for (int k = 0; k < nl; k++) {
        const int i = live[(size_t)k];
        const muse_partial *r = h_recs + (size_t)k * (size_t)top_n;
        double *sc = scores + (size_t)i * (size_t)top_n;
        int64_t *lg = lags + (size_t)i * (size_t)top_n, *ix = series_idx + (size_t)i * (size_t)top_n;
        ctx->multi_rescored += (int64_t)h_cnt[4 * k + 2];
        ctx->multi_refined += (int64_t)h_cnt[4 * k + 3];
        if (r[0].flags == 2) {
            // the list did not fit the device-side select (or the re-scoring launch): this query finishes alone, from its
            // bounds and cut-off on the device (the fused run skips the screening of a prescreened batch)
            muse_batch *b = lb[(size_t)k];
            b->prescreened = 1;
            rc = muse_batch_run_ex(b, nullptr, 0, max_lag, top_n, threshold, sign_filter, MUSE_MODE_SCREEN, 0, sc, lg, ix, n_out + i);
            if (rc) return cleanup(rc);
            continue;
        }
        int64_t c = 0;
        for (; c < top_n && r[c].flags == 0; c++) {
            sc[c] = r[c].score;
            lg[c] = r[c].lag;
            ix[c] = r[c].series_idx;
        }
        n_out[i] = c;
    }
#undef CUM
    return cleanup(MUSE_OK);
}

extern "C" int muse_multi_last_stats(const muse_ctx *ctx, int64_t *n_refined, int64_t *n_rescored) {
    if (!ctx || !n_refined || !n_rescored) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_last_stats: NULL argument");
    *n_refined = ctx->multi_refined;
    *n_rescored = ctx->multi_rescored;
    return MUSE_OK;
}

extern "C" int muse_multi_last_timing(const muse_ctx *ctx, float *ms4) {
    if (!ctx || !ms4) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_last_timing: NULL argument");
    for (int i = 0; i < 4; i++) {
        ms4[i] = 0.f;
        if (ctx->mt_ev[i] && ctx->mt_ev[i + 1] && cudaEventElapsedTime(&ms4[i], ctx->mt_ev[i], ctx->mt_ev[i + 1]) != cudaSuccess) {
            (void)cudaGetLastError();
            ms4[i] = 0.f;
        }
    }
    return MUSE_OK;
}

extern "C" int muse_multi_bounds_tc(muse_ctx *ctx, muse_group *g, const double *refs, int64_t n_refs, int64_t ref_len, float *upper) {
    if (!ctx || !g || !refs || !upper) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_bounds_tc: NULL argument");
    if (n_refs < 1 || n_refs > TcCfg::TN) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_bounds_tc: 1 .. %d references", TcCfg::TN);
    if (!tc_shape_ok(g, ref_len)) return fail(MUSE_ERR_UNSUPPORTED, "the tensor-core bounds exist for even series lengths of FFT length 2048");
    CU(cudaSetDevice(ctx->device));
    const int64_t S = g->size;
    if (S == 0) return MUSE_OK;
    std::vector<muse_batch *> bs;
    int rc = MUSE_OK;
    for (int64_t q = 0; q < n_refs && rc == MUSE_OK; q++) {
        muse_batch *b = nullptr;
        rc = muse_batch_create(ctx, g, refs + (size_t)q * (size_t)ref_len, ref_len, &b);
        if (rc == MUSE_OK) bs.push_back(b);
    }
    if (rc == MUSE_OK) rc = tc_bounds_queue(ctx, g, bs.data(), (int)bs.size());
    for (size_t q = 0; q < bs.size() && rc == MUSE_OK; q++)
        if (cudaMemcpyAsync(upper + q * (size_t)S, bs[q]->d_U, sizeof(float) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            rc = fail(MUSE_ERR_CUDA, "muse_multi_bounds_tc: copy failed");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == MUSE_OK) {
        (void)cudaGetLastError();
        rc = fail(MUSE_ERR_CUDA, "muse_multi_bounds_tc: %s", cudaGetErrorString(cudaPeekAtLastError()));
    }
    for (muse_batch *b : bs) muse_batch_destroy(b);
    return rc;
}

extern "C" int muse_multi_run(muse_ctx *ctx, muse_group *g, const double *refs, int64_t n_refs, int64_t ref_len,
                              const int32_t *key_cols, int32_t n_key_cols, int64_t max_lag, int64_t top_n, double threshold,
                              int32_t sign_filter, int32_t mode, double *scores, int64_t *lags, int64_t *series_idx,
                              int64_t *n_out) {
    if (!ctx || !g || (!refs && n_refs > 0) || !n_out) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_run: NULL argument");
    if (n_refs < 0 || top_n < 0) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_run: n_refs %lld, top_n %lld", (long long)n_refs, (long long)top_n);
    if (top_n > 0 && (!scores || !lags || !series_idx)) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_run: NULL output");
    if (n_key_cols < 0 || (n_key_cols > 0 && !key_cols)) return fail(MUSE_ERR_INVALID_ARG, "muse_multi_run: bad key columns");
    CU(cudaSetDevice(ctx->device));
    // the one-pass multi-query kernel exists for the shape the single-query warp kernel serves (FFT length 2048,
    // ungrouped, unsigned scores, a sign filter unsigned scores can pass, a store worth screening, a device-side
    // top-N); everything else is n_refs x (NewBatch + Run) on the resident store
    const bool one_pass = n_key_cols == 0 && mode != MUSE_MODE_EXACT && sign_filter != MUSE_SIGN_NEG && ref_len == g->N &&
                          next_pow2(ref_len) == 2048 && top_n > 0 && top_n <= 65536 &&
                          top_n * 4 <= g->size && (g->size >= 16384 || mode == MUSE_MODE_SCREEN) && n_refs > 1;
    // the bounds of a launch's queries: one bf16 contraction on the tensor cores for up to 256 queries at a time
    // (MUSE_MULTI_TC=0: the fp32 kernel, 16 queries per pass over the store)
    ctx->multi_refined = ctx->multi_rescored = 0;
    const char *tc_env = getenv("MUSE_MULTI_TC");
    const bool tensor_cores = one_pass && !(tc_env && atoi(tc_env) == 0);
    const int QC = tensor_cores ? MUSE_MULTI_GROUP : ScreenMultiCfg::QC;
    std::vector<muse_batch *> bs((size_t)QC), made((size_t)QC);
    std::vector<int64_t> which((size_t)QC);
    for (int64_t q0 = 0; q0 < n_refs; q0 += QC) {
        const int nq = (int)std::min<int64_t>(QC, n_refs - q0);
        if (tensor_cores && nq > 1) {
            int rct = multi_run_tc_group(ctx, g, refs + (size_t)q0 * (size_t)ref_len, nq, ref_len, max_lag, top_n, threshold, sign_filter,
                                         scores + (size_t)q0 * (size_t)top_n, lags + (size_t)q0 * (size_t)top_n,
                                         series_idx + (size_t)q0 * (size_t)top_n, n_out + q0);
            if (rct) return rct;
            continue;
        }
        int live = 0, rc = MUSE_OK;
        int n_made = 0;
        // the references of the launch: ONE upload (rows at the slab's pitch, pad columns zero), then queued back to
        // back, ONE round trip
        if (ref_len == g->N && (ctx->d_multi_ld != g->ld || ctx->d_multi_n != g->N)) {
            cudaFree(ctx->d_multi_refs);
            ctx->d_multi_refs = nullptr;
            ctx->d_multi_ld = ctx->d_multi_n = 0;
            CU(cudaMalloc(&ctx->d_multi_refs, sizeof(double) * (size_t)MUSE_MULTI_GROUP * (size_t)g->ld));
            CU(cudaMemsetAsync(ctx->d_multi_refs, 0, sizeof(double) * (size_t)MUSE_MULTI_GROUP * (size_t)g->ld, ctx->stream));
            ctx->d_multi_ld = g->ld;
            ctx->d_multi_n = g->N;
        }
        const bool rows_on_device = ref_len == g->N;
        if (rows_on_device)
            CU(cudaMemcpy2DAsync(ctx->d_multi_refs, sizeof(double) * (size_t)g->ld, refs + (size_t)q0 * (size_t)ref_len,
                                 sizeof(double) * (size_t)ref_len, sizeof(double) * (size_t)ref_len, (size_t)nq,
                                 cudaMemcpyHostToDevice, ctx->stream));
        for (int i = 0; i < nq && rc == MUSE_OK; i++) {
            rc = batch_create_queue(ctx, g, refs + (size_t)(q0 + i) * (size_t)ref_len, ref_len, &made[n_made],
                                    rows_on_device ? ctx->d_multi_refs + (size_t)i * (size_t)g->ld : nullptr);
            if (rc == MUSE_OK) n_made++;
        }
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == MUSE_OK) rc = fail(MUSE_ERR_CUDA, "cudaStreamSynchronize failed");
        for (int i = 0; i < n_made; i++) {
            if (rc != MUSE_OK) {
                muse_batch_destroy(made[i]);
                continue;
            }
            const int rf = batch_create_finish(made[i]);
            if (rf == MUSE_ERR_STDDEV_ZERO) {   // muse_batch.go:38-41: this query has no Batch; the others do
                n_out[q0 + i] = -1;
            } else if (rf != MUSE_OK) {
                rc = rf;
            } else {
                bs[live] = made[i];
                which[live++] = q0 + i;
            }
        }
        bool queued = false;
        if (rc == MUSE_OK && one_pass && live > 1) {
            rc = screen_multi(ctx, bs.data(), live, max_lag, top_n, threshold, tensor_cores);
            // the tails of the queries are independent and small (a few thousand series each): every batch queues its
            // tail on its own stream behind the multi-query kernel, so they share the GPU instead of taking turns
            cudaEvent_t screened = nullptr;
            cudaError_t e = cudaEventCreateWithFlags(&screened, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(screened, ctx->stream);
            for (int i = 0; i < live && rc == MUSE_OK && e == cudaSuccess; i++) {
                muse_batch *b = bs[i];
                if (!b->aux) e = cudaStreamCreateWithFlags(&b->aux, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(b->aux, screened, 0);
                if (e != cudaSuccess) break;
                b->use_aux = 1;
                RunArgs a{nullptr, 0, max_lag, top_n, threshold, sign_filter, MUSE_MODE_SCREEN, 0};
                rc = device_topn_queue(b, a);
            }
            for (int i = 0; i < live; i++)
                if (bs[i]->use_aux && cudaStreamSynchronize(bs[i]->aux) != cudaSuccess && e == cudaSuccess) e = cudaErrorUnknown;
            if (screened) cudaEventDestroy(screened);
            if (e != cudaSuccess && rc == MUSE_OK) rc = fail(MUSE_ERR_CUDA, "muse_multi_run: %s", cudaGetErrorString(e));
            queued = rc == MUSE_OK;
        }
        for (int i = 0; i < live; i++) {
            const int64_t q = which[i];
            if (rc == MUSE_OK && queued) {
                RunArgs a{nullptr, 0, max_lag, top_n, threshold, sign_filter, MUSE_MODE_SCREEN, 0};
                n_out[q] = 0;
                rc = device_topn_finish(bs[i], a, scores + (size_t)q * (size_t)top_n, lags + (size_t)q * (size_t)top_n,
                                        series_idx + (size_t)q * (size_t)top_n, n_out + q);
                const unsigned long long *h_n = reinterpret_cast<const unsigned long long *>(bs[i]->h_pin);
                ctx->multi_rescored += (int64_t)h_n[2];
                ctx->multi_refined += (int64_t)h_n[3];
            } else if (rc == MUSE_OK)
                rc = muse_batch_run_ex(bs[i], key_cols, n_key_cols, max_lag, top_n, threshold, sign_filter,
                                       bs[i]->prescreened ? MUSE_MODE_SCREEN : mode, 0,
                                       scores ? scores + (size_t)q * (size_t)top_n : nullptr, lags ? lags + (size_t)q * (size_t)top_n : nullptr,
                                       series_idx ? series_idx + (size_t)q * (size_t)top_n : nullptr, n_out + q);
            muse_batch_destroy(bs[i]);
        }
        if (rc) return rc;
    }
    return MUSE_OK;
}

// ------------------------------------------------------------------------------------
// generic xCorr (xcorr.go:102-153): any n, optional z-normalisation
// ------------------------------------------------------------------------------------
extern "C" int muse_xcorr(muse_ctx *ctx, const double *x, int64_t x_len, const double *y, int64_t y_len, int64_t n,
                          int32_t normalize, double *cc, int64_t cc_capacity, int64_t *n_out, int64_t *lag, double *value,
                          int32_t *std_zero) {
    if (!ctx || !x || !y || !n_out || !lag || !value) return fail(MUSE_ERR_INVALID_ARG, "muse_xcorr: NULL argument");
    if (x_len < 1 || y_len < 1) return fail(MUSE_ERR_INVALID_ARG, "muse_xcorr: empty input (x %lld, y %lld samples)", (long long)x_len, (long long)y_len);
    const int64_t nn = std::max(n, std::max(x_len, y_len));   // xcorr.go:104-106
    // up to 4096 lags: direct evaluation, n^2 fp64 FMAs (the reference's own vectors are n = 5).  Above: the FFT passes of
    // kernels_long.cu -- one transform of length n when n is a power of two (the reference's benchmark shape, n = 32768),
    // else the linear correlation at a power of two >= 2n, folded
    if (nn > (1ll << 26)) return fail(MUSE_ERR_UNSUPPORTED, "muse_xcorr: n = %lld above 2^26", (long long)nn);
    const bool by_fft = nn > 4096;
    long long L = nn;
    if (by_fft && (nn & (nn - 1))) L = next_pow2(2 * nn);
    const size_t work_bytes = by_fft ? long_xcorr_work_bytes(L) : 0;
    if (cc && cc_capacity < nn) return fail(MUSE_ERR_INVALID_ARG, "muse_xcorr: cc holds %lld values, n = %lld", (long long)cc_capacity, (long long)nn);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    *n_out = nn;
    *lag = 0;
    *value = 0.0;
    if (std_zero) *std_zero = 0;
    // one allocation: x, y, xp, yp, cc (doubles), then lag, value, flag
    const size_t nd = (size_t)x_len + (size_t)y_len + 3 * (size_t)nn;
    unsigned char *d = nullptr;
    CU(cudaMalloc(&d, nd * sizeof(double) + 32 + 256 + work_bytes));
    void *d_work = d + ((nd * sizeof(double) + 32 + 255) / 256) * 256;
    double *dx = reinterpret_cast<double *>(d), *dy = dx + x_len, *dxp = dy + y_len, *dyp = dxp + nn, *dcc = dyp + nn;
    long long *dlag = reinterpret_cast<long long *>(dcc + nn);
    double *dval = reinterpret_cast<double *>(dlag + 1);
    int *dflag = reinterpret_cast<int *>(dval + 1);
    struct Tail { long long lag; double val; int flag; int pad; } tail;
    auto body = [&]() -> int {
        CU(cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)x_len, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dy, y, sizeof(double) * (size_t)y_len, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(dlag, 0, 32, st));
        xcorr_prepare_kernel<<<2, 256, 0, st>>>(dx, x_len, dy, y_len, nn, normalize ? 1 : 0, dxp, dyp, dflag);
        // xcorr.go:139-142: gonum's inverse transform is unnormalised (n x the correlation); the reference scales
        // by 1/(n(n-1)) for z-normalised inputs and by 1/n otherwise -> 1/(n-1) and 1 on a direct sum
        const double scale = normalize ? 1.0 / (double)(nn - 1) : 1.0;
        if (by_fft) CU(launch_long_xcorr(dxp, dyp, nn, L, scale, dcc, d_work, st));
        else xcorr_direct_kernel<<<(unsigned)((nn + XC_LAGS - 1) / XC_LAGS), XC_LAGS, 0, st>>>(dxp, dyp, nn, scale, dcc);
        xcorr_argmax_kernel<<<1, 256, 0, st>>>(dcc, nn, dlag, dval);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&tail, dlag, sizeof(tail), cudaMemcpyDeviceToHost, st));
        if (cc) CU(cudaMemcpyAsync(cc, dcc, sizeof(double) * (size_t)nn, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return MUSE_OK;
    };
    const int rc = body();
    cudaFree(d);
    if (rc) return rc;
    if (normalize && tail.flag) {   // xcorr.go:109-126: (nil, 0, 0)
        if (std_zero) *std_zero = 1;
        *n_out = 0;
        return MUSE_OK;
    }
    *lag = tail.lag;
    *value = tail.val;
    return MUSE_OK;
}

// ------------------------------------------------------------------------------------
// multi-GPU exchange over NVLink peer memory (one process per GPU)
// ------------------------------------------------------------------------------------
struct muse_exchange {
    muse_ctx *ctx;
    int rank, world;
    int64_t capacity;                 // records per rank per step
    unsigned char *base;              // local buffer: recs[2][world][capacity] then flags[2][world] (IPC-exported)
    size_t recs_bytes;                // bytes of ONE parity's records
    unsigned char *peer_base[MUSE_EXCHANGE_MAX_RANKS];   // every rank's buffer as seen from here (own: base)
    bool opened;
    unsigned long long epoch;
    unsigned *d_done;
    int *d_status;
    unsigned char *h_recs;            // pinned: the merged top_n records of one step, then the status word
    // device-side merge of the gathered records (group max across shards, filter, top-N)
    MergeState merge;
    size_t merge_bytes;               // one allocation behind merge.hkeys
    unsigned long long *d_nlocal;     // grouped runs: representatives of this shard
    PartialRec *d_out;                // [capacity] merged top_n
};

static size_t exchange_bytes(int world, int64_t capacity) {
    return 2 * (size_t)world * (size_t)capacity * sizeof(muse_partial) + 2 * (size_t)world * sizeof(unsigned long long);
}

extern "C" int muse_exchange_create(muse_ctx *ctx, int32_t rank, int32_t world, int64_t capacity, muse_exchange **out) {
    if (!ctx || !out) return fail(MUSE_ERR_INVALID_ARG, "muse_exchange_create: NULL argument");
    if (world < 1 || world > MUSE_EXCHANGE_MAX_RANKS || rank < 0 || rank >= world || capacity < 1)
        return fail(MUSE_ERR_INVALID_ARG, "muse_exchange_create: rank %d of %d, capacity %lld", rank, world, (long long)capacity);
    CU(cudaSetDevice(ctx->device));
    muse_exchange *x = new muse_exchange();
    memset(x, 0, sizeof(*x));
    x->ctx = ctx;
    x->rank = rank;
    x->world = world;
    x->capacity = capacity;
    x->recs_bytes = (size_t)world * (size_t)capacity * sizeof(muse_partial);
    const size_t bytes = exchange_bytes(world, capacity);
    CU(cudaMalloc(&x->base, bytes));
    CU(cudaMemset(x->base, 0, bytes));
    CU(cudaMalloc(&x->d_done, sizeof(unsigned)));
    CU(cudaMemset(x->d_done, 0, sizeof(unsigned)));
    CU(cudaMalloc(&x->d_status, sizeof(int)));
    CU(cudaMemset(x->d_status, 0, sizeof(int)));
    CU(cudaHostAlloc((void **)&x->h_recs, (size_t)capacity * sizeof(muse_partial) + 64, cudaHostAllocDefault));
    {   // merge scratch: hash table of 2 x (records of all shards) slots, candidate arrays, counters, output records
        const long long total = (long long)world * capacity;
        long long slots = 1;
        while (slots < 2 * total) slots <<= 1;
        const size_t tab = sizeof(unsigned long long) * (size_t)slots;
        const size_t bytes_m = 3 * tab + (size_t)total * (8 + 8 + 4) + 64 + (size_t)capacity * sizeof(PartialRec) + 64;
        unsigned char *m = nullptr;
        CU(cudaMalloc(&m, bytes_m));
        CU(cudaMemset(m, 0, bytes_m));
        x->merge_bytes = 3 * tab;
        x->merge.hkeys = reinterpret_cast<unsigned long long *>(m);
        x->merge.gmax = x->merge.hkeys + slots;
        x->merge.gidx = x->merge.gmax + slots;
        x->merge.slots = slots;
        x->merge.ckey = x->merge.gidx + slots;
        x->merge.cidx = reinterpret_cast<long long *>(x->merge.ckey + total);
        x->merge.clag = reinterpret_cast<int32_t *>(x->merge.cidx + total);
        unsigned char *tail = reinterpret_cast<unsigned char *>(x->merge.clag + total);
        tail = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tail) + 31) & ~(uintptr_t)31);
        x->merge.counters = reinterpret_cast<unsigned long long *>(tail);
        x->d_nlocal = x->merge.counters + 4;
        x->d_out = reinterpret_cast<PartialRec *>(tail + 64);
    }
    x->peer_base[rank] = x->base;
    x->opened = (world == 1);
    CU(cudaDeviceSynchronize());
    *out = x;
    return MUSE_OK;
}

extern "C" int muse_exchange_ipc_handle(muse_exchange *x, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!x || !handle64) return fail(MUSE_ERR_INVALID_ARG, "muse_exchange_ipc_handle: NULL argument");
    if (getenv("MUSE_B200_NO_IPC"))      // test hook: behave like a box where CUDA IPC is closed (callers fall back to the all-gather path)
        return fail(MUSE_ERR_UNSUPPORTED, "CUDA IPC disabled by MUSE_B200_NO_IPC");
    CU(cudaSetDevice(x->ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle64, &h, sizeof(h));
    return MUSE_OK;
}

extern "C" int muse_exchange_open_peers(muse_exchange *x, const void *handles) {
    if (!x || !handles) return fail(MUSE_ERR_INVALID_ARG, "muse_exchange_open_peers: NULL argument");
    CU(cudaSetDevice(x->ctx->device));
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_base[r] = (unsigned char *)p;
    }
    x->opened = true;
    return MUSE_OK;
}

extern "C" void muse_exchange_destroy(muse_exchange *x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->stream);
    for (int r = 0; r < x->world; r++)
        if (r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    cudaFree(x->base);
    cudaFree(x->d_done);
    cudaFree(x->d_status);
    cudaFree(x->merge.hkeys);
    if (x->h_recs) cudaFreeHost(x->h_recs);
    delete x;
}

// One multi-GPU step: scores, then the shard's records stored straight into every rank's receive buffer over NVLink peer
// memory by the kernel that produces them -- ungrouped: the filtered local top_n (partial_topn_push_kernel); grouped: every
// group representative, unfiltered (group_records_push_kernel) --, wait for all the peers' flags, merge ON THE
// DEVICE (group max across shards, filter, top-N) and copy only the top_n records to the host.  Every rank returns the same
// global result.  MUSE_ERR_UNSUPPORTED ("host path needed") when a shard's list was too long for the device-side select or
// for its slot: every rank sees the same marker and can fall back together.
extern "C" int muse_batch_run_exchange_ex(muse_batch *b, muse_exchange *x, const int32_t *key_cols, int32_t n_key_cols, int64_t max_lag,
                                          int64_t top_n, double threshold, int32_t sign_filter, int32_t mode, double *scores, int64_t *lags,
                                          int64_t *series_idx, int64_t *n_out) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!x || !n_out) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_exchange: NULL argument");
    if (x->ctx != b->ctx) return fail(MUSE_ERR_INVALID_ARG, "exchange belongs to another context");
    if (!x->opened) return fail(MUSE_ERR_INVALID_ARG, "muse_exchange_open_peers has not been called");
    if (n_key_cols < 0 || (n_key_cols > 0 && !key_cols)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_exchange: bad key columns");
    if (top_n < 0) top_n = 0;
    if (top_n > x->capacity) return fail(MUSE_ERR_INVALID_ARG, "exchange capacity %lld < top_n %lld", (long long)x->capacity, (long long)top_n);
    if (top_n > 0 && (!scores || !lags || !series_idx)) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_run_exchange: NULL output");
    CU(cudaSetDevice(b->ctx->device));
    *n_out = 0;
    const bool grouped = n_key_cols > 0;
    RunArgs a{key_cols, n_key_cols, max_lag, top_n, threshold, sign_filter, mode, 0};
    a.list_only = grouped ? 0 : 1;
    rc = run_scores(b, a);
    if (rc) return rc;
    const int64_t S = b->g->size;
    cudaStream_t st = bstream(b);
    x->epoch++;
    const int par = (int)(x->epoch & 1ull);
    ExchangePeers ex;
    memset(&ex, 0, sizeof(ex));
    for (int r = 0; r < x->world; r++) {
        ex.recs[r] = reinterpret_cast<PartialRec *>(x->peer_base[r] + (size_t)par * x->recs_bytes);
        ex.flags[r] = reinterpret_cast<unsigned long long *>(x->peer_base[r] + 2 * x->recs_bytes) + (size_t)par * x->world;
    }
    ex.world = x->world;
    ex.rank = x->rank;
    ex.epoch = x->epoch;
    ex.done_blocks = x->d_done;
    if (grouped) {
        // group max and representative of THIS shard on the exact scores (as run_select does), then the push
        GroupTable gt;
        memset(&gt, 0, sizeof(gt));
        KeyCols kc;
        memset(&kc, 0, sizeof(kc));
        int rank_independent = 0;
        rc = key_bits(b->g, key_cols, n_key_cols, kc.bits, &rank_independent);
        if (rc) return rc;
        if (!rank_independent)
            return fail(MUSE_ERR_UNSUPPORTED, "shards merge on group keys of at most %d bits per column (64 / %d keys)", 64 / n_key_cols, n_key_cols);
        rc = setup_group_table(b, a, kc, gt);
        if (rc) return rc;
        const unsigned blocks = (unsigned)((std::max<int64_t>(S, 1) + 255) / 256);
        if (S > 0) {
            group_max_kernel<<<blocks, 256, 0, st>>>(gt, kc, b->d_score, S, b->d_slot);
            group_rep_kernel<<<blocks, 256, 0, st>>>(gt, b->d_score, S, b->d_slot);
        }
        group_records_push_kernel<<<blocks, 256, 0, st>>>(gt, kc, b->d_slot, b->d_score, b->d_lag, (long long)S, (long long)b->g->global_offset,
                                                          (long long)x->capacity, x->d_nlocal, ex);
        b->timing.n_launches += 3;
    } else {
        // counters[0..1] (candidates, selected) are still zero: run_scores cleared all four and ungrouped scoring writes only [2..3]
        if (S > 0) {
            FilterArgs f{a.max_lag, a.threshold, a.sign_filter, 1};
            Cand cand{b->d_ckey, b->d_cidx, b->d_clag, b->d_counters};
            if (b->fused_run == 1) {
                const int64_t lim = std::min<int64_t>(S, MUSE_EXACT_UB);
                emit_listed_kernel<<<(unsigned)((lim + 255) / 256), 256, 0, st>>>(b->d_list, b->d_counters + 2, lim, b->d_score, b->d_lag, f, cand);
            } else {
                GroupTable gt;
                memset(&gt, 0, sizeof(gt));
                emit_candidates_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(gt, nullptr, b->d_score, b->d_lag, S, f, cand);
            }
            b->timing.n_launches++;
        }
        const long long exact_bound = b->fused_run == 1 ? (long long)std::min<int64_t>(S, MUSE_EXACT_UB) : -1;
        const unsigned pblocks = (unsigned)std::min<int64_t>((std::max<int64_t>(S, 1) + 7) / 8, (int64_t)b->ctx->sm_count * 8);
        partial_topn_push_kernel<<<pblocks, 256, 0, st>>>(b->d_ckey, b->d_cidx, b->d_clag, b->d_counters, (long long)top_n,
                                                          (long long)b->g->global_offset, exact_bound, (long long)x->capacity, ex);
        b->timing.n_launches++;
    }
    // ~10 s at ~2 GHz: ranks may be a whole host->device upload apart, but a peer that never arrives must not hang the box
    exchange_wait_kernel<<<1, 32, 0, st>>>(ex.flags[x->rank], x->world, x->epoch, 20000000000ll, x->d_status);
    // merge on the device: the hash table, candidate counter and status start from zero
    CU(cudaMemsetAsync(x->merge.hkeys, 0, x->merge_bytes / 3 * 2, st));                       // keys and maxima
    CU(cudaMemsetAsync(x->merge.gidx, 0xff, x->merge_bytes / 3, st));                         // representatives: +inf
    CU(cudaMemsetAsync(x->merge.counters, 0, sizeof(unsigned long long) * 4, st));
    const long long total = (long long)x->world * x->capacity;
    const unsigned mblocks = (unsigned)((total + 255) / 256);
    const PartialRec *recs_local = ex.recs[x->rank];
    FilterArgs f{a.max_lag, a.threshold, a.sign_filter, 1};
    merge_max_kernel<<<mblocks, 256, 0, st>>>(x->merge, recs_local, ex.flags[x->rank], x->world, (long long)x->capacity);
    merge_rep_kernel<<<mblocks, 256, 0, st>>>(x->merge, recs_local, ex.flags[x->rank], x->world, (long long)x->capacity);
    merge_emit_kernel<<<mblocks, 256, 0, st>>>(x->merge, recs_local, ex.flags[x->rank], x->world, (long long)x->capacity, f);
    merged_topn_kernel<<<(unsigned)std::min<long long>((total + 7) / 8, (long long)b->ctx->sm_count * 8), 256, 0, st>>>(x->merge, (long long)top_n, x->d_out);
    b->timing.n_launches += 5;
    CU(cudaGetLastError());
    int *h_status = reinterpret_cast<int *>(x->h_recs + (size_t)x->capacity * sizeof(muse_partial));
    if (top_n > 0) CU(cudaMemcpyAsync(x->h_recs, x->d_out, sizeof(muse_partial) * (size_t)top_n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_status, x->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(b->h_pin, b->d_counters, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(b->ev[3], st));
    b->timing_pending = 1;
    CU(cudaStreamSynchronize(st));
    if (*h_status) {
        CU(cudaMemsetAsync(x->d_status, 0, sizeof(int), st));
        return fail(MUSE_ERR_CUDA, "muse_batch_run_exchange: a peer did not deliver its records within the time limit");
    }
    const muse_partial *recs = reinterpret_cast<const muse_partial *>(x->h_recs);
    if (top_n > 0 && recs[0].flags == 2)
        return fail(MUSE_ERR_UNSUPPORTED, "host path needed: a shard's record list was too long for the device-side select or merge");
    int64_t k = 0;
    for (; k < top_n && recs[k].flags == 0; k++) {
        scores[k] = recs[k].score;
        lags[k] = recs[k].lag;
        series_idx[k] = recs[k].series_idx;
    }
    *n_out = k;
    return MUSE_OK;
}

extern "C" int muse_batch_run_exchange(muse_batch *b, muse_exchange *x, int64_t max_lag, int64_t top_n, double threshold,
                                       int32_t sign_filter, int32_t mode, double *scores, int64_t *lags, int64_t *series_idx,
                                       int64_t *n_out) {
    return muse_batch_run_exchange_ex(b, x, nullptr, 0, max_lag, top_n, threshold, sign_filter, mode, scores, lags, series_idx, n_out);
}

extern "C" int muse_batch_last_timing(const muse_batch *cb, muse_timing *out) {
    if (!cb || !out) return fail(MUSE_ERR_INVALID_ARG, "muse_batch_last_timing: NULL argument");
    muse_batch *b = const_cast<muse_batch *>(cb);
    if (b->timing_pending) {     // a device-side run (muse_batch_run_partial_device): finish its bookkeeping now
        CU(cudaSetDevice(b->ctx->device));
        CU(cudaEventSynchronize(b->ev[3]));
        read_timing_events(b);
        const unsigned long long *h_n = reinterpret_cast<const unsigned long long *>(b->h_pin);
        if (b->fused_run) {
            b->timing.n_rescored = (int64_t)h_n[2];
            b->timing.n_refined = (int64_t)h_n[3];
        } else {
            b->timing.n_rescored = (int64_t)(h_n[2] + h_n[3]);
        }
        b->timing_pending = 0;
    }
    *out = b->timing;
    return MUSE_OK;
}
