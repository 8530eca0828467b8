// tvf_scene.cuh -- one trial of the synthetic sweep, generated where it is consumed:
// auxiliar_functions/generateSyntheticScene.m:75-111 (N+100 points, projection, Gaussian noise,
// inside-image rejection loop) followed by the column sub-sampling of experiments.m:94-95.
//
// RNG = "TVF scene RNG v1" (tft_vs_fund_b200/scene.py): rng(seed)/rand is MT19937 genrand_res53;
// randn is NumPy's legacy polar method on the same stream; randsample(n,k) is the first k entries of
// NumPy's legacy shuffle of 0..n-1 on a freshly seeded stream.  Integer work (MT19937, rejection masks,
// compaction order, permutation) and the projection arithmetic (fixed order, no FMA contraction, IEEE
// division) are bit-exact with the host generator; the Gaussian uses log(), whose device implementation
// may differ from glibc in the last ulp, so noisy coordinates agree to ~1e-12 px, not bit for bit.
// Host/device code: tests/hostcheck compiles it for the CPU and compares with NumPy.
#pragma once
#include <stdint.h>

#include "tvf_math.cuh"

namespace tvf {

#if defined(__CUDA_ARCH__)
#define TVF_MUL(a, b) __dmul_rn((a), (b))
#define TVF_ADD(a, b) __dadd_rn((a), (b))
#define TVF_DIV(a, b) __ddiv_rn((a), (b))
#define TVF_SQRT(a) __dsqrt_rn((a))
#else
#define TVF_MUL(a, b) ((a) * (b))
#define TVF_ADD(a, b) ((a) + (b))
#define TVF_DIV(a, b) ((a) / (b))
#define TVF_SQRT(a) sqrt((a))
#endif

constexpr int SCENE_MAX_POINTS = 160;       // N + 100 <= 160, i.e. n <= 60 per problem on this path

struct MT19937 {
    uint32_t mt[624];
    int pos;
    int has_gauss;
    double gauss;

    TVF_HD void seed(uint32_t s) {              // numpy mt19937_seed == init_genrand
        for (int i = 0; i < 624; ++i) {
            mt[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624; has_gauss = 0; gauss = 0.0;
    }
    TVF_HD void twist() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
        int i = 0;
        for (; i < 624 - 397; ++i) {
            const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
            mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        }
        for (; i < 623; ++i) {
            const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
            mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        }
        const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        pos = 0;
    }
    TVF_HD uint32_t next32() {
        if (pos == 624) twist();
        uint32_t y = mt[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    TVF_HD double res53() {                      // genrand_res53 == MATLAB rand / RandomState.random_sample
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    TVF_HD double normal() {                     // numpy legacy_gauss (polar method, second value cached)
        if (has_gauss) { has_gauss = 0; const double g = gauss; gauss = 0.0; return g; }
        double x1, x2, r2;
        do {
            x1 = TVF_ADD(TVF_MUL(2.0, res53()), -1.0);
            x2 = TVF_ADD(TVF_MUL(2.0, res53()), -1.0);
            r2 = TVF_ADD(TVF_MUL(x1, x1), TVF_MUL(x2, x2));
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = TVF_SQRT(TVF_DIV(TVF_MUL(-2.0, log(r2)), r2));
        gauss = TVF_MUL(f, x1); has_gauss = 1;
        return TVF_MUL(f, x2);
    }
    TVF_HD uint32_t interval(uint32_t max) {     // numpy random_interval (masked rejection), max <= 2^32-1
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {}
        return v;
    }
};

// P: three 3x4 cameras, ROW-major [view][row][col] (exactly the doubles of scene.scene_cameras);
// out: 6 x n column-major (48 bytes per kept point).  N = n + 100 <= SCENE_MAX_POINTS.
TVF_HD void scene_trial(MT19937& rng, const double* P, int n, double noise, uint32_t seed, double hi_x, double hi_y,
                        double* out, double* c /* 6*N scratch */, unsigned char* arr /* N */, signed char* outpos /* N */) {
    const int N = n + 100;
    // ---- experiments.m:94-95: rng(it); Corresp(:, randsample(N+100, N)) -> first n of a legacy shuffle
    rng.seed(seed);
    for (int i = 0; i < N; ++i) { arr[i] = (unsigned char)i; outpos[i] = -1; }
    for (int i = N - 1; i >= 1; --i) {
        const uint32_t j = rng.interval((uint32_t)i);
        const unsigned char tmp = arr[i]; arr[i] = arr[j]; arr[j] = tmp;
    }
    for (int k = 0; k < n; ++k) outpos[arr[k]] = (signed char)k;
    // ---- generateSyntheticScene.m:75-111
    rng.seed(seed);
    int filled = 0, M = N;
    while (M > 0) {
        for (int i = 0; i < M; ++i) {                                     // X=400*rand(3,M)-200, projected (:82-87)
            const double X = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
            const double Y = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
            const double Z = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
            for (int v = 0; v < 3; ++v) {
                const double* Pv = P + 12 * v;
                double x[3];
                for (int r = 0; r < 3; ++r)
                    x[r] = TVF_ADD(TVF_ADD(TVF_ADD(TVF_MUL(Pv[4 * r], X), TVF_MUL(Pv[4 * r + 1], Y)), TVF_MUL(Pv[4 * r + 2], Z)), Pv[4 * r + 3]);
                c[6 * i + 2 * v] = TVF_DIV(x[0], x[2]);
                c[6 * i + 2 * v + 1] = TVF_DIV(x[1], x[2]);
            }
        }
        for (int v = 0; v < 3; ++v)                                       // x_noise = x + randn(2,M)*noise (:90-92)
            for (int i = 0; i < M; ++i) {
                const double z0 = rng.normal(), z1 = rng.normal();
                c[6 * i + 2 * v] = TVF_ADD(c[6 * i + 2 * v], TVF_MUL(z0, noise));
                c[6 * i + 2 * v + 1] = TVF_ADD(c[6 * i + 2 * v + 1], TVF_MUL(z1, noise));
            }
        for (int i = 0; i < M; ++i) {                                     // inside image, compaction in order (:95-107)
            bool inside = true;
            for (int v = 0; v < 3; ++v) {
                const double x = c[6 * i + 2 * v], y = c[6 * i + 2 * v + 1];
                inside = inside && (x <= hi_x) && (y <= hi_y) && (x >= 0.0) && (y >= 0.0);
            }
            if (inside) {
                const int k = outpos[filled++];
                if (k >= 0)
                    for (int q = 0; q < 6; ++q) out[6 * k + q] = c[6 * i + q];
            }
        }
        M = N - filled;                                                   // :110
    }
}

}  // namespace tvf
