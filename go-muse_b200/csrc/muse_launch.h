// muse_launch.h -- host-side launchers of the templated score kernels.  Each family of instantiations is its own
// translation unit (kernels_*.cu) so that the library builds in parallel; muse_api.cu sees only these entry points.
#pragma once

#include <cuda_runtime.h>

#include "muse_exact.cuh"
#include "muse_screen_big.cuh"
#include "muse_screen_wide.cuh"
#include "muse_screen_block.cuh"
#include "muse_screen_sub.cuh"
#include "muse_screen_multi.cuh"
#include "muse_bounds_tc.cuh"

namespace muse {

// score_exact_kernel<log2m, ., mode> (muse_exact.cuh); log2m = log2(n/2) in 0 .. 13
cudaError_t launch_exact(int mode, int log2m, const ExactParams &p, cudaStream_t st);
cudaError_t launch_exact_batch(int mode, const ExactParams *d_table, int nq, int64_t max_count, cudaStream_t st);
// points per thread of the exact kernel for each FFT size (fixes the twiddle layout of ExactParams::twM)
inline int exact_log2p(int log2m) { return log2m < 4 ? log2m : 4; }

// fp32 screening + fused second stage: warp kernel (n = 2048), block kernel (n = 512, 1024, 4096 .. 16384)
cudaError_t launch_screen_warp(const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_block(int log2m, const ScreenParams &p, int sm_count, cudaStream_t st);
// n = 128 .. 1024: sixteen .. two series per warp (muse_screen_sub.cuh)
cudaError_t launch_screen_sub1(const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_sub2(const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_sub3(const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_sub4(const ScreenParams &p, int sm_count, cudaStream_t st);
// n = 4096 .. 16384 (muse_screen_big.cuh)
cudaError_t launch_screen_big(int log2m, const ScreenParams &p, int sm_count, cudaStream_t st);
// n = 16384: nz = rows of 256 complex slots that hold samples, 17 .. 32, four per translation unit
cudaError_t launch_screen_big13_a(int nz, const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_big13_b(int nz, const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_big13_c(int nz, const ScreenParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_screen_big13_d(int nz, const ScreenParams &p, int sm_count, cudaStream_t st);
// n = 16384 at twice the occupancy (muse_screen_wide.cuh); its twiddle tables are fill_wide_twiddles'
cudaError_t launch_screen_wide(const ScreenParams &p, int sm_count, cudaStream_t st);
// the same pass for up to ScreenMultiCfg::QC reference queries at once (n = 2048)
cudaError_t launch_screen_multi(const ScreenParams &p, const MultiQuery *d_queries, int nq, int sm_count, cudaStream_t st);

// the second stage for up to RefineMultiCfg::QMAX queries with precomputed bounds (n = 2048)
cudaError_t launch_refine_multi(const ScreenParams &p, const MultiQuery *d_queries, int nq, int sm_count, unsigned *d_next, cudaStream_t st);
// many queries' bounds as one bf16 contraction on the tensor cores (muse_bounds_tc.cuh)
cudaError_t launch_mag_tiles(const ScreenParams &p, unsigned char *a_tiles, float *mid, int sm_count, cudaStream_t st);
cudaError_t launch_weight_tiles(const float4 *const *d_sw, int nq, unsigned char *b_tiles, cudaStream_t st);
cudaError_t launch_bounds_tc(const TcBoundsParams &p, cudaStream_t st);

// ---- series above n = 16384 (kernels_long.cu): Stockham passes through global memory, two series per transform ----
struct LongParams {
    const double *slab;      // [rows][ld] fp64
    int64_t ld;
    int64_t count;           // series to process
    const int32_t *idx;      // optional gather list
    int N;                   // series length
    int log2n;               // FFT length n = 2^log2n
    int signed_scores;
    const cd *X;             // FFT_n of the reference row / ((N-1) std n), n entries
    const cd *tw;            // exp(-2*pi*i*k/n), n entries
    double *out_score;       // MODE_SCORE: [rows]   MODE_CC: cc[n]
    int32_t *out_lag;
    cd *out_X;               // MODE_REF
    int32_t *out_flag;       // MODE_REF / MODE_CC: 1 when std == 0
    double *stats;           // set by launch_long (inside the work buffer)
};
size_t long_work_bytes(int log2n, long long chunk_pairs);
cudaError_t launch_long_twiddles(cd *tw, int log2n, cudaStream_t st);
cudaError_t launch_long(int mode, LongParams p, void *work, long long chunk_pairs, cudaStream_t st);
// generic xCorr through the same passes: L = nn (a power of two) or a power of two >= 2 nn
size_t long_xcorr_work_bytes(long long L);
cudaError_t launch_long_xcorr(const double *xp, const double *yp, long long nn, long long L, double scale, double *cc, void *work,
                              cudaStream_t st);

}  // namespace muse
