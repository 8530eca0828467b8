// tvf_warp.cuh -- warp-cooperative pieces: shuffles, reductions and the
// "smallest eigenvector of a small SPD matrix" solver used for the 27x27
// trilinearity Gram, its 15x15 projection and the 9x9 eight-point Gram.
//
// Storage: lane r owns row r of the symmetric matrix in registers g[0..N).
// A right-looking Cholesky runs in place; because the trailing matrix stays
// symmetric, lane k's entries g[m], m>k, are dead after step k and are reused
// to hold column k of L.  After the factorisation lane j therefore holds
//   g[m] = L(j,m) for m<=j   and   g[m] = L(m,j) for m>j,
// which is exactly what the forward (L y = b) and backward (L' x = y)
// substitutions need in their axpy form: one broadcast shuffle + one DFMA per
// step, no cross-lane reduction.  The eigenvector is obtained by inverse
// iteration with a tiny relative diagonal shift (keeps the factorisation
// positive for noise-free, exactly singular data).
#pragma once
#include "tvf_math.cuh"

namespace tvf {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// g: lane's row (lanes >= N must pass zeros).  Returns this lane's component of
// the unit eigenvector of the smallest eigenvalue; *converged tells whether the
// iteration met its tolerance (warp-uniform).
template <int N>
__device__ __forceinline__ double smallest_eigvec_spd(double (&g)[N], const int lane, bool* converged) {
    // relative shift: delta = 1e-13 * trace/N
    double diag = 0.0;
#pragma unroll
    for (int m = 0; m < N; ++m) diag = (lane == m) ? g[m] : diag;
    const double tr = warp_sum(diag);
    const double delta = 1.0e-13 * tr * (1.0 / N);
    const double floor_piv = 1.0e-3 * delta + 1e-300;
#pragma unroll
    for (int m = 0; m < N; ++m) g[m] += (lane == m) ? delta : 0.0;

    double dinv = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double dk = shfl_d(g[k], k);
        dk = fmax(dk, floor_piv);
        const double rinv = rsqrt(dk);
        double lk = g[k] * rinv;
        if (lane == k) { lk = dk * rinv; dinv = rinv; }
        if (lane >= k) g[k] = lk;
        const double lkz = (lane > k) ? lk : 0.0;
#pragma unroll
        for (int m = k + 1; m < N; ++m) {
            const double lm = shfl_d(lk, m);
            const double upd = fma(-lkz, lm, g[m]);
            g[m] = (lane == k) ? lm : upd;
        }
    }

    // x0: solve L' x = 1
    double x = 0.0;
    {
        double acc = (lane < N) ? 1.0 : 0.0;
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            const double xj = shfl_d(acc * dinv, j);
            acc = fma(-g[j], xj, acc);
            x = (lane == j) ? xj : x;
        }
        x *= rsqrt(warp_sum(x * x));
    }
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < 60; ++it) {
        double acc = x, y = 0.0, z = 0.0;
#pragma unroll
        for (int m = 0; m < N; ++m) {                 // L y = x
            const double ym = shfl_d(acc * dinv, m);
            acc = fma(-g[m], ym, acc);
            y = (lane == m) ? ym : y;
        }
        acc = y;
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {            // L' z = y
            const double zj = shfl_d(acc * dinv, j);
            acc = fma(-g[j], zj, acc);
            z = (lane == j) ? zj : z;
        }
        z *= rsqrt(warp_sum(z * z));
        const double diff = warp_max(fabs(z - x));
        x = z;
        if (!(diff > 1.0e-15)) { ok = true; break; }
    }
    *converged = ok;
    return x;
}

}  // namespace tvf
