// muse_synth.cuh -- siggen-style synthetic series (rect / line / noise), counter based so
// that host and device produce bit-identical doubles without a host array ever
// existing (806 GB cannot be staged through the host).
//
// Stands in for go-matrixprofile's siggen.Rect / Line / Noise / Add as used by
// example_test.go:16-47 and muse_batch_test.go:137-146 (test data only in the
// reference).  Only integer hashing, exact int->double conversion and single fp64
// add/mul/fma operations are used, so there is no room for host/device divergence.
#pragma once

#include <math.h>
#include <stdint.h>

#include "muse_fft.cuh"

namespace muse {

// a*b rounded once and never contracted into an FMA with a following add
MUSE_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}

MUSE_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform in [0, 1) keyed by (seed, i, t)
MUSE_HD double synth_u01(uint64_t seed, uint64_t i, uint64_t t) {
    const uint64_t h = mix64(mix64(seed ^ (i * 0xD1B54A32D192ED03ull)) + t);
    return mul_rn((double)(h >> 11), 1.0 / 9007199254740992.0);
}

struct SynthSeries {
    int kind;        // 0 rect+noise, 1 line+noise, 2 noise
    double amp;      // rect amplitude
    int64_t start, end;   // rect support [start, end)
    double slope, offset;
};

// variant 0: the benchmark's mix (kind = i mod 3, rect widths 3 .. 20); variant 1: every series a rect of the reference's
// width (10 samples) at a random position -- the adversarial store for the screening: nearly every score lies within the
// slack of the top-N cut-off
MUSE_HD SynthSeries synth_params(uint64_t seed, int64_t i, int64_t N, int variant = 0) {
    SynthSeries s;
    const uint64_t ui = (uint64_t)i;
    s.kind = variant == 1 ? 0 : (int)(ui % 3ull);
    const uint64_t P = 0xFFFFFFFF00000000ull;   // parameter stream: t values no sample uses
    const double u0 = synth_u01(seed, ui, P + 0), u1 = synth_u01(seed, ui, P + 1), u2 = synth_u01(seed, ui, P + 2);
    s.amp = fma(39.5, u0, 0.5);                                    // U(0.5, 40)
    const int64_t half = N / 8;                                   // mid ~ N/2 +- N/8
    const int64_t mid = N / 2 - half + (int64_t)mul_rn(u1, (double)(2 * half + 1));
    const int64_t width = variant == 1 ? 10 : 3 + (int64_t)mul_rn(u2, 18.0);               // U{3..20}
    s.start = mid - width / 2;
    s.end = s.start + width;
    s.slope = fma(0.02, u0, -0.01);                                 // U(-0.01, 0.01)
    s.offset = u1;
    return s;
}

MUSE_HD double synth_value(uint64_t seed, int64_t i, const SynthSeries &s, int64_t t) {
    const double noise = mul_rn(0.1, synth_u01(seed, (uint64_t)i, (uint64_t)t) - 0.5);   // siggen.Noise(0.1, n)
    if (s.kind == 0) return (t >= s.start && t < s.end) ? noise + s.amp : noise;   // Rect + Noise
    if (s.kind == 1) return fma(s.slope, (double)t, s.offset) + noise;              // Line + Noise
    return noise;
}

// The reference query: Rect(1.5, N/2, 10) + Noise(0.1) (example_test.go:15-20 scaled to N).
MUSE_HD double synth_ref_value(uint64_t seed, int64_t N, int64_t t) {
    const double noise = mul_rn(0.1, synth_u01(seed, 0xFFFFFFFFFFFFFFFEull, (uint64_t)t) - 0.5);
    const int64_t start = N / 2 - 5;
    return (t >= start && t < start + 10) ? noise + 1.5 : noise;
}

}  // namespace muse
