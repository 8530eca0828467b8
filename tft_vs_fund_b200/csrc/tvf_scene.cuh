// tvf_scene.cuh -- one trial of the synthetic sweep, generated where it is consumed:
// auxiliar_functions/generateSyntheticScene.m:75-111 (N+100 points, projection, Gaussian noise,
// inside-image rejection loop) followed by the column sub-sampling of experiments.m:94-95.
//
// RNG = "TVF scene RNG v2" (tft_vs_fund_b200/scene.py): rng(seed)/rand is MT19937 genrand_res53;
// randn is the polar (Marsaglia) method on the same stream in NumPy-legacy order (second value of a pair
// first, the other cached) with the logarithm taken by tvf_log below -- a fixed sequence of IEEE
// operations, so that host and device produce the same bits (libm's log differs between glibc and CUDA in
// the last ulp); randsample(n,k) is the first k entries of NumPy's legacy shuffle of 0..n-1 on a freshly
// seeded stream.  Everything -- MT19937, rejection masks, compaction order, permutation, projections (fixed
// order, no FMA contraction, IEEE division) and the noisy coordinates -- is bit-exact with the host generator.
// Host/device code: tests/hostcheck compiles it for the CPU and compares with NumPy.
#pragma once
#include <stdint.h>
#include <string.h>

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

// Natural logarithm of a positive, normal double as a fixed sequence of correctly rounded IEEE operations
// (no FMA, no libm): x = m * 2^e with m in (sqrt(1/2), sqrt(2)], s = (m-1)/(m+1), z = s^2,
// log m = 2s + 2s*z*(1/3 + z/5 + ... + z^10/23) (next term < 6e-19 relative), log x = e*ln2_hi + (log m + e*ln2_lo).
// Accurate to ~2 ulp; its point is reproducibility: scene.py / oracle/scene.py restate it in NumPy.
TVF_HD double tvf_log(double x) {
    long long bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(x);
#else
    memcpy(&bits, &x, 8);
#endif
    long long e = ((bits >> 52) & 0x7ff) - 1023;
    long long mb = (bits & 0x000fffffffffffffLL) | (1023LL << 52);
    if (mb > 0x3ff6a09e667f3bcdLL) { mb -= (1LL << 52); e += 1; }        // m > sqrt(2): halve
    double m;
#if defined(__CUDA_ARCH__)
    m = __longlong_as_double(mb);
#else
    memcpy(&m, &mb, 8);
#endif
    const double f = TVF_ADD(m, -1.0);
    const double s = TVF_DIV(f, TVF_ADD(2.0, f));
    const double z = TVF_MUL(s, s);
    double q = 1.0 / 23.0;
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 21.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 19.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 17.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 15.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 13.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 11.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 9.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 7.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 5.0);
    q = TVF_ADD(TVF_MUL(q, z), 1.0 / 3.0);
    const double t = TVF_MUL(2.0, s);
    const double lg = TVF_ADD(t, TVF_MUL(TVF_MUL(t, z), q));
    const double ed = (double)e;
    return TVF_ADD(TVF_MUL(ed, 6.93147180369123816490e-01), TVF_ADD(lg, TVF_MUL(ed, 1.90821492927058770002e-10)));
}

constexpr int SCENE_MAX_POINTS = 160;       // N + 100 <= 160, i.e. n <= 60 per problem on this path

struct MT19937 {
    uint32_t mt[624];
    int pos;
    int has_gauss;
    int twists;          // state regenerations since seed(): lets a caller rewind cheaply when none happened
    double gauss;

    TVF_HD void seed(uint32_t s) {              // numpy mt19937_seed == init_genrand
        for (int i = 0; i < 624; ++i) {
            mt[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624; has_gauss = 0; gauss = 0.0; twists = 0;
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
        pos = 0; ++twists;
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
        const double f = TVF_SQRT(TVF_DIV(TVF_MUL(-2.0, tvf_log(r2)), r2));
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

// All L noise levels of ONE seed (experiments.m:91-95: the trials j = (seed-1)*L + lv share rng(seed)).  The first
// pass of generateSyntheticScene's while-loop (:80-92) does not depend on the noise level -- same 3-D points, same
// projections, same Gaussian draws, same number of variates consumed -- and neither does the sub-sample
// permutation, so both are computed once; per level only the scaling by `noise`, the inside-image mask and the
// (short) refill passes differ.  The refill passes continue from the stream position the first pass left: the
// generator is rewound to a snapshot (just the read position unless a state regeneration happened in between).
// Produces exactly the bits of scene_trial for every level.  Levels [lv_lo, lv_hi) are written to
// out + (lv - lv_lo)*6*n.  Scratch: clean, z, c of 6*N doubles each; arr, outpos of N.
TVF_HD void scene_seed_levels(MT19937& rng, MT19937& snap, const double* P, int n, const double* noise_levels, int lv_lo,
                              int lv_hi, uint32_t seed, double hi_x, double hi_y, double* out, double* clean, double* z,
                              double* c, unsigned char* arr, signed char* outpos) {
    const int N = n + 100;
    rng.seed(seed);
    for (int i = 0; i < N; ++i) { arr[i] = (unsigned char)i; outpos[i] = -1; }
    for (int i = N - 1; i >= 1; --i) {
        const uint32_t j = rng.interval((uint32_t)i);
        const unsigned char tmp = arr[i]; arr[i] = arr[j]; arr[j] = tmp;
    }
    for (int k = 0; k < n; ++k) outpos[arr[k]] = (signed char)k;
    rng.seed(seed);
    for (int i = 0; i < N; ++i) {                                         // first pass, noise-independent part
        const double X = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
        const double Y = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
        const double Z = TVF_ADD(TVF_MUL(400.0, rng.res53()), -200.0);
        for (int v = 0; v < 3; ++v) {
            const double* Pv = P + 12 * v;
            double x[3];
            for (int r = 0; r < 3; ++r)
                x[r] = TVF_ADD(TVF_ADD(TVF_ADD(TVF_MUL(Pv[4 * r], X), TVF_MUL(Pv[4 * r + 1], Y)), TVF_MUL(Pv[4 * r + 2], Z)), Pv[4 * r + 3]);
            clean[6 * i + 2 * v] = TVF_DIV(x[0], x[2]);
            clean[6 * i + 2 * v + 1] = TVF_DIV(x[1], x[2]);
        }
    }
    for (int v = 0; v < 3; ++v)
        for (int i = 0; i < N; ++i) { z[6 * i + 2 * v] = rng.normal(); z[6 * i + 2 * v + 1] = rng.normal(); }
    snap = rng;
    for (int lv = lv_lo; lv < lv_hi; ++lv) {
        const double noise = noise_levels[lv];
        double* o = out + (size_t)(lv - lv_lo) * 6 * n;
        if (rng.twists != snap.twists) rng = snap;
        else { rng.pos = snap.pos; rng.has_gauss = snap.has_gauss; rng.gauss = snap.gauss; }
        int filled = 0;
        for (int i = 0; i < N; ++i) {
            double p[6];
            bool inside = true;
            for (int v = 0; v < 3; ++v) {
                const double x = TVF_ADD(clean[6 * i + 2 * v], TVF_MUL(z[6 * i + 2 * v], noise));
                const double y = TVF_ADD(clean[6 * i + 2 * v + 1], TVF_MUL(z[6 * i + 2 * v + 1], noise));
                inside = inside && (x <= hi_x) && (y <= hi_y) && (x >= 0.0) && (y >= 0.0);
                p[2 * v] = x; p[2 * v + 1] = y;
            }
            if (inside) {
                const int k = outpos[filled++];
                if (k >= 0)
                    for (int q = 0; q < 6; ++q) o[6 * k + q] = p[q];
            }
        }
        int M = N - filled;
        while (M > 0) {                                                   // refill passes: as in scene_trial
            for (int i = 0; i < M; ++i) {
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
            for (int v = 0; v < 3; ++v)
                for (int i = 0; i < M; ++i) {
                    const double z0 = rng.normal(), z1 = rng.normal();
                    c[6 * i + 2 * v] = TVF_ADD(c[6 * i + 2 * v], TVF_MUL(z0, noise));
                    c[6 * i + 2 * v + 1] = TVF_ADD(c[6 * i + 2 * v + 1], TVF_MUL(z1, noise));
                }
            for (int i = 0; i < M; ++i) {
                bool inside = true;
                for (int v = 0; v < 3; ++v) {
                    const double x = c[6 * i + 2 * v], y = c[6 * i + 2 * v + 1];
                    inside = inside && (x <= hi_x) && (y <= hi_y) && (x >= 0.0) && (y >= 0.0);
                }
                if (inside) {
                    const int k = outpos[filled++];
                    if (k >= 0)
                        for (int q = 0; q < 6; ++q) o[6 * k + q] = c[6 * i + q];
                }
            }
            M = N - filled;
        }
    }
}

}  // namespace tvf
