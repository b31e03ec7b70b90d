// kernels_screen_big13_b.cu -- score_screen_big_kernel<13, ., NZ> (muse_screen_big.cuh) for NZ = 21 .. 24 rows of samples.
#include <algorithm>

#include "muse_launch.h"

namespace muse {

#ifndef MUSE_BIG_MINB13
#define MUSE_BIG_MINB13 2
#endif

template <int NZ>
static cudaError_t launch_big13_nz(const ScreenParams &p, int sm_count, cudaStream_t st) {
    using C = ScreenBigCfg<13>;
    auto kern = score_screen_big_kernel<13, MUSE_BIG_MINB13, NZ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, C::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t blocks = std::min<int64_t>(p.count, (int64_t)sm_count * per_sm);      // persistent: contiguous ranges of series
    kern<<<(unsigned)blocks, C::T, C::SMEM, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_screen_big13_b(int nz, const ScreenParams &p, int sm_count, cudaStream_t st) {
    switch (nz) {
        case 21: return launch_big13_nz<21>(p, sm_count, st);
        case 22: return launch_big13_nz<22>(p, sm_count, st);
        case 23: return launch_big13_nz<23>(p, sm_count, st);
        case 24: return launch_big13_nz<24>(p, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace muse
