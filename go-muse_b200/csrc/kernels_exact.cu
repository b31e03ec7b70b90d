// kernels_exact.cu -- instantiations of score_exact_kernel (muse_exact.cuh) for every FFT length n = 2 .. 16384.
#include <cstring>

#include "muse_launch.h"

namespace muse {

template <int LOG2M, int LOG2P, int MODE, int MINB>
static cudaError_t launch_exact_cfg(const ExactParams &p, cudaStream_t st) {
    using C = ExactCfg<LOG2M, LOG2P>;
    auto kern = score_exact_kernel<LOG2M, LOG2P, MODE, MINB>;
    if (C::SMEM > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (p.count + C::SPB - 1) / C::SPB;
    kern<<<(unsigned)blocks, C::TB, C::SMEM, st>>>(p);
    return cudaGetLastError();
}

template <int LOG2M, int MODE>
static cudaError_t launch_exact_t(const ExactParams &p, cudaStream_t st) {
    constexpr int LOG2P = LOG2M < 4 ? LOG2M : 4;
    // 512 resident threads per SM = a 128-register cap: measured best on B200 at n = 2048
    // (16 warps/SM, 12.3 ms per 1M series vs 15.5 ms uncapped at 8 warps/SM; profiles/r01_tune_exact.txt)
    constexpr int MINB = 512 / ExactCfg<LOG2M, LOG2P>::TB > 0 ? 512 / ExactCfg<LOG2M, LOG2P>::TB : 1;
    return launch_exact_cfg<LOG2M, LOG2P, MODE, MINB>(p, st);
}

template <int MODE>
static cudaError_t launch_exact_m(int log2m, const ExactParams &p, cudaStream_t st) {
    switch (log2m) {
#define MUSE_CASE(L) case L: return launch_exact_t<L, MODE>(p, st);
        MUSE_CASE(0) MUSE_CASE(1) MUSE_CASE(2) MUSE_CASE(3) MUSE_CASE(4) MUSE_CASE(5) MUSE_CASE(6)
        MUSE_CASE(7) MUSE_CASE(8) MUSE_CASE(9) MUSE_CASE(10) MUSE_CASE(11) MUSE_CASE(12) MUSE_CASE(13)
#undef MUSE_CASE
    }
    return cudaErrorInvalidValue;
}

// n = 2048 only (the multi-query path): d_table[nq] parameter blocks in device memory, max_count = the largest count
cudaError_t launch_exact_batch(int mode, const ExactParams *d_table, int nq, int64_t max_count, cudaStream_t st) {
    constexpr int LOG2M = 10, LOG2P = 4;
    using C = ExactCfg<LOG2M, LOG2P>;
    constexpr int MINB = 512 / C::TB > 0 ? 512 / C::TB : 1;
    ExactParams p;
    memset(&p, 0, sizeof(p));
    p.batch = d_table;
    const dim3 grid((unsigned)((max_count + C::SPB - 1) / C::SPB), (unsigned)nq);
    if (mode == MODE_SCORE) score_exact_kernel<LOG2M, LOG2P, MODE_SCORE, MINB, true><<<grid, C::TB, C::SMEM, st>>>(p);
    else if (mode == MODE_REF) score_exact_kernel<LOG2M, LOG2P, MODE_REF, MINB, true><<<grid, C::TB, C::SMEM, st>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_exact(int mode, int log2m, const ExactParams &p, cudaStream_t st) {
    switch (mode) {
        case MODE_SCORE: return launch_exact_m<MODE_SCORE>(log2m, p, st);
        case MODE_REF: return launch_exact_m<MODE_REF>(log2m, p, st);
        case MODE_CC: return launch_exact_m<MODE_CC>(log2m, p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace muse
