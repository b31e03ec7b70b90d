// kernels_screen_block.cu -- instantiations of score_screen_block_kernel (muse_screen_block.cuh).
#include <algorithm>

#include "muse_launch.h"

namespace muse {

template <int LOG2M, int MINB>
static cudaError_t launch_screen_block_t(const ScreenParams &p, int sm_count, cudaStream_t st) {
    using C = ScreenBlockCfg<LOG2M>;
    auto kern = score_screen_block_kernel<LOG2M, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, C::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t blocks = std::min<int64_t>(p.count, (int64_t)sm_count * per_sm);     // persistent: grid-stride over the series
    kern<<<(unsigned)blocks, C::T, C::SMEM, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_screen_block(int log2m, const ScreenParams &p, int sm_count, cudaStream_t st) {
    switch (log2m) {
        // n = 4096 .. 16384 moved to muse_screen_big.cuh, n = 512 / 1024 to muse_screen_sub.cuh: what is left here is round 1's
        // kernel for A/B runs of those two sizes (MUSE_BLOCK_SMALL=1)
        case 9: return launch_screen_block_t<9, 16>(p, sm_count, st);
        case 8: return launch_screen_block_t<8, 16>(p, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace muse
