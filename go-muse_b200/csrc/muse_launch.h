// muse_launch.h -- host-side launchers of the templated score kernels.  Each family of instantiations is its own
// translation unit (kernels_*.cu) so that the library builds in parallel; muse_api.cu sees only these entry points.
#pragma once

#include <cuda_runtime.h>

#include "muse_exact.cuh"
#include "muse_screen_big.cuh"
#include "muse_screen_wide.cuh"
#include "muse_screen_block.cuh"
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
// n = 4096 .. 16384 (muse_screen_big.cuh)
cudaError_t launch_screen_big(int log2m, const ScreenParams &p, int sm_count, cudaStream_t st);
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

}  // namespace muse
