/*
 * synth_gen.c -- the benchmark's synthetic siggen-style rows (rect / line / noise) on the host.
 * TEST INFRASTRUCTURE ONLY (bench.py's CPU arms and the tests): the same counter-based generator as
 * go-muse_b200/csrc/muse_synth.cuh, restated in C so that the reference arm of bench.py needs nothing but oracle/.
 * Stands in for go-matrixprofile's siggen.Rect / Line / Noise / Add as used by example_test.go:16-47 and
 * muse_batch_test.go:137-146 (test data only in the reference).  Compiled with -ffp-contract=off: every operation
 * below is a single correctly rounded IEEE operation (or an explicit fma), so host and device agree bit for bit
 * (tests/test_oracle_c.py checks it against the library's generator).
 */
#include <math.h>
#include <stdint.h>

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double u01(uint64_t seed, uint64_t i, uint64_t t) {
    const uint64_t h = mix64(mix64(seed ^ (i * 0xD1B54A32D192ED03ull)) + t);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

void muse_oracle_synth_row(uint64_t seed, int64_t index, int64_t N, double *out) {
    const uint64_t ui = (uint64_t)index;
    const int kind = (int)(ui % 3ull);
    const uint64_t P = 0xFFFFFFFF00000000ull;
    const double u0 = u01(seed, ui, P + 0), u1 = u01(seed, ui, P + 1), u2 = u01(seed, ui, P + 2);
    const double amp = fma(39.5, u0, 0.5);
    const int64_t half = N / 8;
    const int64_t mid = N / 2 - half + (int64_t)(u1 * (double)(2 * half + 1));
    const int64_t width = 3 + (int64_t)(u2 * 18.0);
    const int64_t start = mid - width / 2, end = start + width;
    const double slope = fma(0.02, u0, -0.01), offset = u1;
    for (int64_t t = 0; t < N; t++) {
        const double noise = 0.1 * (u01(seed, ui, (uint64_t)t) - 0.5);
        if (kind == 0) out[t] = (t >= start && t < end) ? noise + amp : noise;
        else if (kind == 1) out[t] = fma(slope, (double)t, offset) + noise;
        else out[t] = noise;
    }
}

void muse_oracle_synth_rows(uint64_t seed, int64_t first, int64_t count, int64_t N, double *out) {
    for (int64_t r = 0; r < count; r++) muse_oracle_synth_row(seed, first + r, N, out + r * N);
}

void muse_oracle_synth_reference(uint64_t seed, int64_t N, double *out) {
    const int64_t start = N / 2 - 5;
    for (int64_t t = 0; t < N; t++) {
        const double noise = 0.1 * (u01(seed, 0xFFFFFFFFFFFFFFFEull, (uint64_t)t) - 0.5);
        out[t] = (t >= start && t < start + 10) ? noise + 1.5 : noise;
    }
}
