// tvf_warp.cuh -- warp-cooperative pieces: shuffles, reductions and the
// "smallest eigenvector of a small SPD matrix" solver used for the 27x27
// trilinearity Gram, its 15x15 projection and the 9x9 eight-point Gram.
//
// Storage: lane r owns row r of the symmetric matrix in registers g[0..N).
// The matrix is inverted in place by N Gauss-Jordan sweeps (no pivoting: SPD),
// fully unrolled so every register index is static.  Per sweep k each lane
// publishes its scaled column entry A(j,k)/d to shared memory (by symmetry
// that vector is row k), reads the row back as broadcast 128-bit loads and does
// one DFMA per matrix element:  A(j,m) -= c_j * A(k,m)/d.
// The pivot lane needs A(k,m) <- A(k,m)/d instead; with the matrix pre-scaled
// to unit trace every pivot d <= 1 and that is the *same* DFMA with
// c_k = d - 1 (no cancellation: A + (1-d)*A/d), so the special case costs one
// scalar select per sweep instead of two per element.
// The inverse then drives power iteration (= inverse iteration on G): one
// broadcast mat-vec of N DFMAs per step.  A relative diagonal shift of 1e-13
// keeps noise-free, exactly singular systems factorable; it does not change
// the eigenvectors.  tests/kernel_model.py restates this algorithm in NumPy.
#pragma once
#include "tvf_math.cuh"

namespace tvf {

constexpr unsigned FULL = 0xffffffffu;
constexpr int EIG_MAX_ITER = 80;
#ifndef TVF_EIG_PIVOTNORM
#define TVF_EIG_PIVOTNORM 1
#endif
// software-pipelined Gauss-Jordan sweeps: measured 4.6 % SLOWER in stage 1 (profiles/r01_variants.md), off
#ifndef TVF_GJ_PIPE
#define TVF_GJ_PIPE 0
#endif
#ifndef TVF_EIG_TOL
#define TVF_EIG_TOL 4.0e-15
#endif
constexpr double EIG_TOL = TVF_EIG_TOL;      // largest component change of the unit eigenvector between two steps

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// 1/d for a normal, positive d: MUFU.RCP64H seed (>= 20 bits) + two Newton steps (full double precision,
// not correctly rounded) -- 7 instructions instead of the ~20 of the IEEE division with its slow path.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

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

// g: this lane's row (lanes >= N must pass zeros).  sbuf: 64 doubles of shared memory private to
// the warp, 16-byte aligned.  Returns this lane's component of the unit eigenvector belonging to
// the smallest eigenvalue; *converged is warp-uniform.
struct NoRefine {
    __device__ __forceinline__ double operator()(const double*) const { return 0.0; }
};

// Transposed butterfly over 32 per-lane values: afterwards lane L holds the warp-wide total of v[L].
__device__ __forceinline__ double warp_reduce_transposed32(double (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool hi = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = hi ? v[i] : v[i + half];
            const double keep = hi ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, half);
        }
    }
    return v[0];
}

// `resid(xbuf)`: optional functor returning this lane's component of (A'A) x for the x stored in shared
// memory at xbuf[0..N), computed from the UN-squared design rows.  With nrefine > 0 the eigenvector is then
// polished by iterative refinement, x <- x - (G+dI)^-1 (A'(A x) - rho x): the residual carries an error of
// eps*|A|*|Ax| instead of the eps*|A|^2 of the Gram route, which matters for barely determined systems
// (n = 7..11) where sigma_{N-1} is tiny.  nrefine is warp-uniform.
template <int N, class Resid = NoRefine>
__device__ __forceinline__ double smallest_eigvec_spd(double (&g)[N], const int lane, double* sbuf, bool* converged,
                                                      Resid resid = Resid(), int nrefine = 0) {
    static_assert(N <= 32 && N >= 2, "one matrix row per lane");
    constexpr int NP = (N + 1) & ~1;          // row length padded to a whole number of 128-bit loads
    double diag = 0.0;
#pragma unroll
    for (int m = 0; m < N; ++m) diag = (lane == m) ? g[m] : diag;
    const double tr = warp_sum(diag);
    const double sc = 1.0 / tr;
    const double delta = 1.0e-13 / N;         // relative to the unit trace
    const double floor_piv = 1.0e-3 * delta;
#pragma unroll
    for (int m = 0; m < N; ++m) g[m] = g[m] * sc + ((lane == m) ? delta : 0.0);

    __syncwarp();                              // sbuf may still be read by a previous phase
    if (lane >= N && lane < NP) { sbuf[lane] = 0.0; sbuf[32 + lane] = 0.0; }
#if TVF_GJ_PIPE
    // Software-pipelined sweeps (same operations, same results bit for bit; REJECTED by measurement, kept as a knob): the
    // serial part of a sweep -- pivot
    // broadcast, reciprocal (five dependent FP64 operations), scaled column, its publication and the barrier before
    // the row can be read back -- is ~110 cycles in which a warp has nothing else to issue.  Sweep k therefore updates
    // column k+1 FIRST and starts that chain for sweep k+1 at once (into the other row buffer), so it runs under the
    // remaining 25 DFMAs of sweep k.
    double colj = g[0];
    double piv = fast_rcp(fmax(shfl_d(colj, 0), floor_piv));
    double rk = colj * piv;
    if (lane < N) sbuf[lane] = rk;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        __syncwarp();                          // row k is in sbuf[(k & 1) * 32 ..]; the other buffer is free again
        const double2* b2 = reinterpret_cast<const double2*>(sbuf + (k & 1) * 32);
        const double c = colj - ((lane == k) ? 1.0 : 0.0);
        double colj_n = 0.0, piv_n = 0.0, rk_n = 0.0;
        if (k + 1 < N) {
            const double2 r = b2[(k + 1) >> 1];
            g[k + 1] = fma(-c, ((k + 1) & 1) ? r.y : r.x, g[k + 1]);
            colj_n = g[k + 1];
            piv_n = fast_rcp(fmax(shfl_d(colj_n, k + 1), floor_piv));
            rk_n = colj_n * piv_n;
            if (lane < N) sbuf[((k + 1) & 1) * 32 + lane] = rk_n;
        }
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            if (2 * m2 != k && 2 * m2 != k + 1 && 2 * m2 < N) g[2 * m2] = fma(-c, r.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 != k + 1 && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-c, r.y, g[2 * m2 + 1]);
        }
        g[k] = (lane == k) ? -piv : rk;
        colj = colj_n; piv = piv_n; rk = rk_n;
    }
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        const double d = fmax(shfl_d(colj, k), floor_piv);
        const double piv = fast_rcp(d);
        const double rk = colj * piv;
        double* buf = sbuf + (k & 1) * 32;
        if (lane < N) buf[lane] = rk;
        __syncwarp();
        const double c = colj - ((lane == k) ? 1.0 : 0.0);
        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            if (2 * m2 != k && 2 * m2 < N) g[2 * m2] = fma(-c, r.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-c, r.y, g[2 * m2 + 1]);
        }
        g[k] = (lane == k) ? -piv : rk;
    }
#endif
    __syncwarp();                              // last sweep's row buffer is reused below
    // g now holds -(G + delta I)^-1 (scaled).  Power iteration on its negative.
    bool ok = false;
#if TVF_EIG_PIVOTNORM
    // The iterate is kept normalised to "largest component = 1": the pivot lane comes from one REDUX over the high
    // words and a ballot, its value from one shuffle, so the serial chain product -> normalise -> vote is a broadcast and
    // a reciprocal instead of a five-level shuffle reduction and a reciprocal square root.  Dividing by the signed pivot
    // also removes the sign flip of -M.  One exact 2-norm normalisation follows the loop.
    double x = (lane < N) ? 1.0 : 0.0;
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* buf = sbuf + (it & 1) * 32;
        if (lane < NP) buf[lane] = x;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(buf);
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            if (2 * m2 < N) z0 = fma(g[2 * m2], r.x, z0);
            if (2 * m2 + 1 < N) z1 = fma(g[2 * m2 + 1], r.y, z1);
        }
        double z = z0 + z1;
        const unsigned hz = (unsigned)__double2hiint(z) & 0x7fffffffu;
        const unsigned hmax = __reduce_max_sync(FULL, hz);
        const int piv = __ffs(__ballot_sync(FULL, hz == hmax)) - 1;
        z *= fast_rcp(shfl_d(z, piv));
        const bool moving = fabs(z - x) > EIG_TOL;            // max_lane |z - x| > tol, as one warp vote
        x = z;
        if (!__any_sync(FULL, moving)) { ok = true; break; }
    }
    x *= rsqrt_(warp_sum(x * x));
#else
    double x = (lane < N) ? rsqrt_((double)N) : 0.0;
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* buf = sbuf + (it & 1) * 32;
        if (lane < NP) buf[lane] = x;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(buf);
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            if (2 * m2 < N) z0 = fma(g[2 * m2], r.x, z0);
            if (2 * m2 + 1 < N) z1 = fma(g[2 * m2 + 1], r.y, z1);
        }
        double z = -(z0 + z1);
        z *= rsqrt_(warp_sum(z * z));
        const bool moving = fabs(z - x) > EIG_TOL;            // max_lane |z - x| > tol, as one warp vote
        x = z;
        if (!__any_sync(FULL, moving)) { ok = true; break; }
    }
#endif
    for (int step = 0; step < nrefine; ++step) {
        __syncwarp();
        if (lane < NP) sbuf[lane] = x;
        __syncwarp();
        const double gl = resid(sbuf) * sc;                  // (G x)_lane of the unit-trace matrix, from the rows
        const double rho = warp_sum(x * gl);
        const double r = (lane < N) ? gl - rho * x : 0.0;
        __syncwarp();
        if (lane < NP) sbuf[32 + lane] = r;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(sbuf + 32);
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 rr = b2[m2];
            if (2 * m2 < N) z0 = fma(g[2 * m2], rr.x, z0);
            if (2 * m2 + 1 < N) z1 = fma(g[2 * m2 + 1], rr.y, z1);
        }
        x += z0 + z1;                                        // g holds -(G+dI)^-1
        x *= rsqrt_(warp_sum(x * x));
    }
    *converged = ok;
    return x;
}

// ---------------------------------------------------------------------------------------------------------------
// Two independent problems per warp: the same solver on HALF-warps (N <= 16; lane = 16 h + r, lane r of half h owns row
// r of problem h).  The 15 x 15 projected system of linearTFT's second step keeps 15 of 32 lanes busy in the full-warp
// solver; here both halves work.  Nothing crosses between the halves except the loop-exit vote: each half freezes its
// iterate at the step where IT converged, so a problem's result does not depend on its neighbour (bit-reproducible under
// any batching).  sbuf: 64 doubles; half h uses [16 h, 16 h + 16) of each 32-double parity buffer.
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <int N>
__device__ __forceinline__ double smallest_eigvec_spd_half(double (&g)[N], const int lane, double* sbuf, bool* converged) {
    static_assert(N <= 16 && N >= 2, "one matrix row per lane of a half-warp");
    const int r = lane & 15, h16 = lane & 16;
    const unsigned hmask = 0xffffu << h16;
    double diag = 0.0;
#pragma unroll
    for (int m = 0; m < N; ++m) diag = (r == m) ? g[m] : diag;
    const double tr = half_sum(diag);
    const double sc = 1.0 / tr;
    const double delta = 1.0e-13 / N;         // relative to the unit trace
    const double floor_piv = 1.0e-3 * delta;
#pragma unroll
    for (int m = 0; m < N; ++m) g[m] = g[m] * sc + ((r == m) ? delta : 0.0);

    __syncwarp();                              // sbuf may still be read by a previous phase
    if (r >= N) { sbuf[lane] = 0.0; sbuf[32 + lane] = 0.0; }
#if TVF_GJ_PIPE
    double colj = g[0];                          // software-pipelined as in smallest_eigvec_spd
    double piv = fast_rcp(fmax(__shfl_sync(FULL, colj, 0, 16), floor_piv));
    double rk = colj * piv;
    if (r < N) sbuf[h16 + r] = rk;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(sbuf + (k & 1) * 32 + h16);
        const double c = colj - ((r == k) ? 1.0 : 0.0);
        double colj_n = 0.0, piv_n = 0.0, rk_n = 0.0;
        if (k + 1 < N) {
            const double2 q = b2[(k + 1) >> 1];
            g[k + 1] = fma(-c, ((k + 1) & 1) ? q.y : q.x, g[k + 1]);
            colj_n = g[k + 1];
            piv_n = fast_rcp(fmax(__shfl_sync(FULL, colj_n, k + 1, 16), floor_piv));
            rk_n = colj_n * piv_n;
            if (r < N) sbuf[((k + 1) & 1) * 32 + h16 + r] = rk_n;
        }
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) {
            const double2 q = b2[m2];
            if (2 * m2 != k && 2 * m2 != k + 1 && 2 * m2 < N) g[2 * m2] = fma(-c, q.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 != k + 1 && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-c, q.y, g[2 * m2 + 1]);
        }
        g[k] = (r == k) ? -piv : rk;
        colj = colj_n; piv = piv_n; rk = rk_n;
    }
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        const double d = fmax(__shfl_sync(FULL, colj, k, 16), floor_piv);
        const double piv = fast_rcp(d);
        const double rk = colj * piv;
        double* buf = sbuf + (k & 1) * 32 + h16;
        if (r < N) buf[r] = rk;
        __syncwarp();
        const double c = colj - ((r == k) ? 1.0 : 0.0);
        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) {
            const double2 q = b2[m2];
            if (2 * m2 != k && 2 * m2 < N) g[2 * m2] = fma(-c, q.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-c, q.y, g[2 * m2 + 1]);
        }
        g[k] = (r == k) ? -piv : rk;
    }
#endif
    __syncwarp();
    // g holds -(G + delta I)^-1.  Pivot-normalised power iteration (see smallest_eigvec_spd), per half.
    double x = (r < N) ? 1.0 : 0.0;
    bool done = false;                          // uniform within a half
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* buf = sbuf + (it & 1) * 32 + h16;
        buf[r] = x;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(buf);
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) {
            const double2 q = b2[m2];
            if (2 * m2 < N) z0 = fma(g[2 * m2], q.x, z0);
            if (2 * m2 + 1 < N) z1 = fma(g[2 * m2 + 1], q.y, z1);
        }
        double z = z0 + z1;
        const unsigned hz = (unsigned)__double2hiint(z) & 0x7fffffffu;
        const unsigned hmax = __reduce_max_sync(hmask, hz);
        const int piv = __ffs(__ballot_sync(hmask, hz == hmax)) - 1;
        z *= fast_rcp(shfl_d(z, piv));
        const bool moving = !done && (fabs(z - x) > EIG_TOL);
        if (!done) x = z;
        const unsigned bal = __ballot_sync(FULL, moving);
        done = done || ((bal & hmask) == 0u);
        if (bal == 0u) break;
    }
    x *= rsqrt_(half_sum(x * x));
    *converged = done;
    return x;
}

}  // namespace tvf
