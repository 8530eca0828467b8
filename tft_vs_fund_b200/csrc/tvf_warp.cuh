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
// Gauss-Jordan sweeps publish the RAW pivot column (one STS straight from the register, before the pivot's reciprocal
// chain starts) and scale the per-lane multiplier instead of the published row: the shared-memory round trip
// (STS -> barrier -> LDS.128) then runs under the shuffle + MUFU + Newton chain instead of after it.
#ifndef TVF_GJ_RAWCOL
#define TVF_GJ_RAWCOL 0
#endif
#ifndef TVF_GJ2_FUSEDRCP
#define TVF_GJ2_FUSEDRCP 1
#endif
#ifndef TVF_PI_STICKY
#define TVF_PI_STICKY 1
#endif
#ifndef TVF_EIG_TOL
#define TVF_EIG_TOL 4.0e-15
#endif
// Start vector of the power iterations: M (e_0 + e_{N/2} + e_{N-1}) instead of the vector of ones -- with M = -(G + delta I)^-1
// in registers row by row, that product is the sum of three entries of the lane's own row: the first application of M costs two
// additions instead of a mat-vec (one iteration of ~7 saved; the fixed point and the stopping test are unchanged).
#ifndef TVF_PI_FREE_START
#define TVF_PI_FREE_START 1
#endif
// Largest change of a component's MAGNITUDE between two steps.  Magnitudes, not signed values: the iterate is normalised
// by its largest component, and when two components of opposite sign tie for that role to within rounding the pivot can
// alternate between them from step to step, flipping the sign of the whole (otherwise converged) vector each time.  The
// sign of the eigenvector is arbitrary anyway (T, F are defined up to sign).
constexpr double EIG_TOL = TVF_EIG_TOL;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// 1/d for a normal, positive d: MUFU.RCP64H seed (>= 20 bits) + two Newton steps (full double precision,
// not correctly rounded) -- 7 instructions instead of the ~20 of the IEEE division with its slow path.
// TVF_RCP_CUBIC: one third-order step r (1 + e + e^2), e = 1 - d r, instead of two Newton steps: the seed's 2^-20 becomes
// 2^-60 in THREE dependent DFMAs instead of four (the reciprocal sits on the serial chain of every Gauss-Jordan sweep).
#ifndef TVF_RCP_CUBIC
#define TVF_RCP_CUBIC 0
#endif
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#if TVF_RCP_CUBIC
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#endif
}

// max(d, floor) for the Gauss-Jordan pivots, decided on the high word with one integer compare (negative, zero and
// tiny pivots all compare below; a NaN pivot passes through and poisons the problem, which is then flagged NONFINITE)
// instead of fmax()'s DSETP.MAX + selects + NaN-quieting sequence.  Identical to fmax for every regular pivot.
#ifndef TVF_PIVOT_INTFLOOR
#define TVF_PIVOT_INTFLOOR 1
#endif
__device__ __forceinline__ double pivot_floor(double d, double floor_piv) {
#if TVF_PIVOT_INTFLOOR
    return (__double2hiint(d) < __double2hiint(floor_piv)) ? floor_piv : d;
#else
    return fmax(d, floor_piv);
#endif
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
// PRESCALED: the caller passes G/trace + delta*I already (the Gram is gathered from a moment table, and scaling that
// table costs 3 multiplications per lane instead of N per lane plus the diagonal-extraction select chain); no refinement.
template <int N, class Resid = NoRefine, bool PRESCALED = false>
__device__ __forceinline__ double smallest_eigvec_spd(double (&g)[N], const int lane, double* sbuf, bool* converged,
                                                      Resid resid = Resid(), int nrefine = 0) {
    static_assert(N <= 32 && N >= 2, "one matrix row per lane");
    constexpr int NP = (N + 1) & ~1;          // row length padded to a whole number of 128-bit loads
    const double delta = 1.0e-13 / N;         // relative to the unit trace
    const double floor_piv = 1.0e-3 * delta;
    double sc = 1.0;
    if (!PRESCALED) {
        double diag = 0.0;
#pragma unroll
        for (int m = 0; m < N; ++m) diag = (lane == m) ? g[m] : diag;
        const double tr = warp_sum(diag);
        sc = 1.0 / tr;
#pragma unroll
        for (int m = 0; m < N; ++m) g[m] = g[m] * sc + ((lane == m) ? delta : 0.0);
    }

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
    double piv = fast_rcp(pivot_floor(shfl_d(colj, 0), floor_piv));
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
            piv_n = fast_rcp(pivot_floor(shfl_d(colj_n, k + 1), floor_piv));
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
#elif TVF_GJ_RAWCOL
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        double* buf = sbuf + (k & 1) * 32;
        if (lane < N) buf[lane] = colj;        // raw column k == raw row k (symmetry)
        const double d = pivot_floor(shfl_d(colj, k), floor_piv);
        const double piv = fast_rcp(d);
        __syncwarp();
        const double cs = (colj - ((lane == k) ? 1.0 : 0.0)) * piv;
        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            if (2 * m2 != k && 2 * m2 < N) g[2 * m2] = fma(-cs, r.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-cs, r.y, g[2 * m2 + 1]);
        }
        g[k] = (lane == k) ? -piv : colj * piv;
    }
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        const double d = pivot_floor(shfl_d(colj, k), floor_piv);
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
#if TVF_PI_STICKY
    int piv = 0;
#endif
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* buf = sbuf + (it & 1) * 32;
        if (lane < NP) buf[lane] = x;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(buf);
        double za[4] = {0.0, 0.0, 0.0, 0.0};               // four independent accumulation chains
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 r = b2[m2];
            const int c = (m2 & 1) * 2;
            if (2 * m2 < N) za[c] = fma(g[2 * m2], r.x, za[c]);
            if (2 * m2 + 1 < N) za[c + 1] = fma(g[2 * m2 + 1], r.y, za[c + 1]);
        }
        double z = (za[0] + za[1]) + (za[2] + za[3]);
        const unsigned hz = (unsigned)__double2hiint(z) & 0x7fffffffu;
        const unsigned hmax = __reduce_max_sync(FULL, hz);
#if TVF_PI_STICKY
        // the pivot lane is kept while its component stays within one binade of the largest: the shuffle runs beside the
        // REDUX instead of behind REDUX -> ballot -> ffs, and two components of (almost) equal magnitude cannot make the
        // normalisation alternate between them (which would never meet the tolerance although the direction converged)
        double pv = shfl_d(z, piv);
        if (((unsigned)__double2hiint(pv) & 0x7fffffffu) + 0x00100000u < hmax) {
            piv = __ffs(__ballot_sync(FULL, hz == hmax)) - 1;
            pv = shfl_d(z, piv);
        }
        z *= fast_rcp(pv);
#else
        const int piv = __ffs(__ballot_sync(FULL, hz == hmax)) - 1;
        z *= fast_rcp(shfl_d(z, piv));
#endif
        const bool moving = fabs(fabs(z) - fabs(x)) > EIG_TOL;            // max_lane |z - x| > tol, as one warp vote
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
        const bool moving = fabs(fabs(z) - fabs(x)) > EIG_TOL;            // max_lane |z - x| > tol, as one warp vote
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
// max over the 16 lanes of a half-warp with FULL-mask shuffles (xor distances < 16 stay inside the half).  Collectives with a
// partial, lane-dependent member mask (__reduce_max_sync(0xffff << h16, ..)) compile to a WARPSYNC.COLLECTIVE loop with
// divergent branches: ncu attributed 8 + 4 % of the two-row solver's samples to `branch_resolving` there.
__device__ __forceinline__ unsigned half_max_u32(unsigned v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// the same maximum as two FULL-mask REDUX operations (one per half, the other half contributing zeros): the two are
// independent, so they pipeline, where the xor ladder is four dependent shuffles
__device__ __forceinline__ unsigned half_max2_u32(unsigned v, int h16) {
    const unsigned lo = __reduce_max_sync(FULL, h16 ? 0u : v), hi = __reduce_max_sync(FULL, h16 ? v : 0u);
    return h16 ? hi : lo;
}

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
    double piv = fast_rcp(pivot_floor(__shfl_sync(FULL, colj, 0, 16), floor_piv));
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
            piv_n = fast_rcp(pivot_floor(__shfl_sync(FULL, colj_n, k + 1, 16), floor_piv));
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
#elif TVF_GJ_RAWCOL
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        double* buf = sbuf + (k & 1) * 32 + h16;
        if (r < N) buf[r] = colj;
        const double d = pivot_floor(__shfl_sync(FULL, colj, k, 16), floor_piv);
        const double piv = fast_rcp(d);
        __syncwarp();
        const double cs = (colj - ((r == k) ? 1.0 : 0.0)) * piv;
        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) {
            const double2 q = b2[m2];
            if (2 * m2 != k && 2 * m2 < N) g[2 * m2] = fma(-cs, q.x, g[2 * m2]);
            if (2 * m2 + 1 != k && 2 * m2 + 1 < N) g[2 * m2 + 1] = fma(-cs, q.y, g[2 * m2 + 1]);
        }
        g[k] = (r == k) ? -piv : colj * piv;
    }
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double colj = g[k];
        const double d = pivot_floor(__shfl_sync(FULL, colj, k, 16), floor_piv);
#if TVF_GJ2_FUSEDRCP
        double seed;                                     // see smallest_eigvec_spd_half2: three dependent operations behind the seed
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(d));
        const double e = fma(-d, seed, 1.0);
        const double u = colj * seed;
        const double t = fma(e, e, e);
        const double rk = fma(u, t, u);
        const double piv = fma(seed, t, seed);
#else
        const double piv = fast_rcp(d);
        const double rk = colj * piv;
#endif
        double* buf = sbuf + (k & 1) * 32 + h16;
        buf[r] = rk;                               // lanes r >= N hold zero rows: they write zeros into padding nobody uses
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
#if TVF_PI_FREE_START
    double x = (r < N) ? (g[0] + g[N / 2]) + g[N - 1] : 0.0;
#else
    double x = (r < N) ? 1.0 : 0.0;
#endif
    bool done = false;                          // uniform within a half
#if TVF_PI_STICKY
    int piv = h16;                              // absolute lane of this half's pivot component
#endif
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
#if TVF_PI_STICKY
        const unsigned hmax = half_max2_u32(hz, h16);
        double pv = shfl_d(z, piv);
        const bool stale = ((unsigned)__double2hiint(pv) & 0x7fffffffu) + 0x00100000u < hmax;     // uniform within a half
        if (__any_sync(FULL, stale)) {
            const int np = __ffs(__ballot_sync(FULL, hz == hmax) & hmask) - 1;
            if (stale) piv = np;
            pv = shfl_d(z, piv);
        }
        z *= fast_rcp(pv);
#else
        const unsigned hmax = half_max_u32(hz);
        const int piv = __ffs(__ballot_sync(FULL, hz == hmax) & hmask) - 1;
        z *= fast_rcp(shfl_d(z, piv));
#endif
        const bool moving = !done && (fabs(fabs(z) - fabs(x)) > EIG_TOL);
        if (!done) x = z;
        const unsigned bal = __ballot_sync(FULL, moving);
        done = done || ((bal & hmask) == 0u);
        if (bal == 0u) break;
    }
    x *= rsqrt_(half_sum(x * x));
    *converged = done;
    return x;
}

// ---------------------------------------------------------------------------------------------------------------
// Two problems per warp for systems of up to 32 unknowns: half h of the warp owns problem h and lane r of a half owns
// the TWO rows r and r + NR of its matrix (NR = ceil(N/2) <= 16; N = 27: rows r and r + 14, lanes 14 and 15 idle).
// Compared with one row per lane on a full warp, a Gauss-Jordan sweep still costs one DFMA per matrix element, but the
// pivot-row read-back (N/2 broadcast 128-bit shared loads), the pivot's reciprocal chain, the publication and the
// barrier are issued once per TWO problems: 31 FP64 + 13 other instructions per problem and sweep instead of 34 + 30.
//
// The caller passes a matrix that is ALREADY scaled to unit trace and shifted (G/tr + delta*I): the 27 x 27 Gram is
// gathered from a moment table, so scaling the table (96 multiplications) replaces scaling the matrix (729) and the
// diagonal-extraction select chain.  Rows >= N (the second row of the last lanes) must be passed as zeros: they are
// never pivots and stay zero.  sbuf: 128 doubles, 16-byte aligned; half h uses [32 h, 32 h + 32) of each 64-double
// parity buffer.  Nothing crosses between the halves except the loop-exit vote (each half freezes its iterate when IT
// has converged), so a problem's result does not depend on its neighbour.
#ifndef TVF_GJ2_PIPE
#define TVF_GJ2_PIPE 0
#endif

template <int N>
__device__ __forceinline__ void smallest_eigvec_spd_half2(double (&g0)[N], double (&g1)[N], const int lane, double* sbuf,
                                                          double* x0_out, double* x1_out, bool* converged) {
    static_assert(N <= 32 && N >= 4, "two matrix rows per lane of a half-warp");
    constexpr int NR = (N + 1) / 2;
    constexpr int NP = (N + 1) & ~1;
    const int r = lane & 15, h32 = (lane & 16) * 2;          // offset of this half's 32-double segment
    const unsigned hmask = 0xffffu << (lane & 16);
    const double floor_piv = 1.0e-3 * (1.0e-13 / N);
    const bool own0 = r < NR, own1 = r + NR < N;
    static_assert(((N + 1) & ~1) <= 30, "slots 30 and 31 of a half's segment take the idle lanes' stores");
    const int slot0 = own0 ? r : 30, slot1 = own1 ? NR + r : 31;
    __syncwarp();                              // sbuf may still be read by a previous phase
    if (r == 0) {                              // padding entry of the row buffers (odd N)
        if (NP > N) { sbuf[h32 + N] = 0.0; sbuf[64 + h32 + N] = 0.0; }
    }
#if TVF_GJ2_PIPE
    // Software-pipelined sweeps, written for ONE or TWO resident warps per scheduler (255 registers): every warp has to keep
    // the FP64 pipe busy on its own.  (1) All N/2 broadcast loads of pivot row k are issued at once, right behind the
    // barrier, so the DFMAs never wait for shared memory one load at a time.  (2) Sweep k updates column k+1 FIRST and
    // then starts sweep k+1's serial chain -- pivot shuffle, reciprocal, scaled column, publication into the other
    // parity buffer -- which completes under the remaining 2(N-2) DFMAs of sweep k.  Same operations on the same
    // operands as the plain loop below: bit-identical results.
    const int dst0 = slot0, dst1 = slot1;                            // idle lanes write two unused slots of the segment
    double c0j = g0[0], c1j = g1[0];
    double piv = fast_rcp(pivot_floor(__shfl_sync(FULL, c0j, 0, 16), floor_piv));
    double r0 = c0j * piv, r1 = c1j * piv;
    sbuf[h32 + dst0] = r0; sbuf[h32 + dst1] = r1;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int a = (k >= NR) ? 1 : 0, rk = k - a * NR;     // pivot row k lives on lane rk of each half, array a
        __syncwarp();                          // row k is in parity buffer k & 1; the other one is free again
        const double2* b2 = reinterpret_cast<const double2*>(sbuf + (k & 1) * 64 + h32);
#if TVF_GJ2_PIPE == 1
        double2 q[NP / 2];
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) q[m2] = b2[m2];
#endif
        const double c0 = c0j - ((a == 0 && r == rk) ? 1.0 : 0.0);
        const double c1 = c1j - ((a == 1 && r == rk) ? 1.0 : 0.0);
        double n0j = 0.0, n1j = 0.0, npiv = 0.0, nr0 = 0.0, nr1 = 0.0;
        if (k + 1 < N) {
            const int an = (k + 1 >= NR) ? 1 : 0, rkn = k + 1 - an * NR;
#if TVF_GJ2_PIPE == 1
            const double v = ((k + 1) & 1) ? q[(k + 1) >> 1].y : q[(k + 1) >> 1].x;
#else
            const double v = sbuf[(k & 1) * 64 + h32 + k + 1];     // variant 2: loads stay one at a time (no 56-register row)
#endif
            g0[k + 1] = fma(-c0, v, g0[k + 1]); g1[k + 1] = fma(-c1, v, g1[k + 1]);
            n0j = g0[k + 1]; n1j = g1[k + 1];
            npiv = fast_rcp(pivot_floor(__shfl_sync(FULL, an ? n1j : n0j, rkn, 16), floor_piv));
            nr0 = n0j * npiv; nr1 = n1j * npiv;
            double* bn = sbuf + ((k + 1) & 1) * 64 + h32;
            bn[dst0] = nr0; bn[dst1] = nr1;
        }
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
#if TVF_GJ2_PIPE == 1
            const double2 qq = q[m2];
#else
            const double2 qq = b2[m2];
#endif
            if (2 * m2 != k && 2 * m2 != k + 1 && 2 * m2 < N) { g0[2 * m2] = fma(-c0, qq.x, g0[2 * m2]); g1[2 * m2] = fma(-c1, qq.x, g1[2 * m2]); }
            if (2 * m2 + 1 != k && 2 * m2 + 1 != k + 1 && 2 * m2 + 1 < N) { g0[2 * m2 + 1] = fma(-c0, qq.y, g0[2 * m2 + 1]); g1[2 * m2 + 1] = fma(-c1, qq.y, g1[2 * m2 + 1]); }
        }
        g0[k] = (a == 0 && r == rk) ? -piv : r0;
        g1[k] = (a == 1 && r == rk) ? -piv : r1;
        c0j = n0j; c1j = n1j; piv = npiv; r0 = nr0; r1 = nr1;
    }
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int a = (k >= NR) ? 1 : 0, rk = k - a * NR;     // pivot row k lives on lane rk of each half, array a
        const double c0j = g0[k], c1j = g1[k];
        const double d = pivot_floor(__shfl_sync(FULL, a ? c1j : c0j, rk, 16), floor_piv);
#if TVF_GJ2_FUSEDRCP
        // the scaled column leaves the reciprocal's refinement in THREE dependent FP64 operations behind the MUFU seed
        // (e | c*seed -> e + e^2 -> c*seed*(1 + e + e^2)) instead of five (two Newton steps, then the product): the seed's
        // 2^-20 relative error becomes e^3 = 2^-60, and this chain is the serial part of every sweep
        double seed;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(d));
        const double e = fma(-d, seed, 1.0);
        const double u0 = c0j * seed, u1 = c1j * seed;
        const double t = fma(e, e, e);
        const double r0 = fma(u0, t, u0), r1 = fma(u1, t, u1);
        const double piv = fma(seed, t, seed);
#else
        const double piv = fast_rcp(d);
        const double r0 = c0j * piv, r1 = c1j * piv;
#endif
        double* buf = sbuf + (k & 1) * 64 + h32;
        buf[slot0] = r0;                           // unconditional stores: idle lanes write the unused slots 30 / 31
        buf[slot1] = r1;
        __syncwarp();
        const double c0 = c0j - ((a == 0 && r == rk) ? 1.0 : 0.0);
        const double c1 = c1j - ((a == 1 && r == rk) ? 1.0 : 0.0);
        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 q = b2[m2];
            if (2 * m2 != k && 2 * m2 < N) { g0[2 * m2] = fma(-c0, q.x, g0[2 * m2]); g1[2 * m2] = fma(-c1, q.x, g1[2 * m2]); }
            if (2 * m2 + 1 != k && 2 * m2 + 1 < N) { g0[2 * m2 + 1] = fma(-c0, q.y, g0[2 * m2 + 1]); g1[2 * m2 + 1] = fma(-c1, q.y, g1[2 * m2 + 1]); }
        }
        g0[k] = (a == 0 && r == rk) ? -piv : r0;
        g1[k] = (a == 1 && r == rk) ? -piv : r1;
    }
#endif
    __syncwarp();
    // g holds -(G + delta I)^-1.  Pivot-normalised power iteration (see smallest_eigvec_spd), per half.
#if TVF_PI_FREE_START
    double x0 = own0 ? (g0[0] + g0[N / 2]) + g0[N - 1] : 0.0, x1 = own1 ? (g1[0] + g1[N / 2]) + g1[N - 1] : 0.0;
#else
    double x0 = own0 ? 1.0 : 0.0, x1 = own1 ? 1.0 : 0.0;
#endif
    bool done = false;                          // uniform within a half
#if TVF_PI_STICKY
    int src = lane & 16;                        // pivot component: row of array 0 (pivA) or 1 on absolute lane src
    bool pivA = true;
#endif
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* buf = sbuf + (it & 1) * 64 + h32;
        buf[slot0] = x0;
        buf[slot1] = x1;
        __syncwarp();
        const double2* b2 = reinterpret_cast<const double2*>(buf);
        // four independent accumulation chains per row (the mat-vec is latency bound: 3 resident warps per scheduler)
        double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m2 = 0; m2 < NP / 2; ++m2) {
            const double2 q = b2[m2];
            const int c = (m2 & 1) * 2;
            if (2 * m2 < N) { a0[c] = fma(g0[2 * m2], q.x, a0[c]); a1[c] = fma(g1[2 * m2], q.x, a1[c]); }
            if (2 * m2 + 1 < N) { a0[c + 1] = fma(g0[2 * m2 + 1], q.y, a0[c + 1]); a1[c + 1] = fma(g1[2 * m2 + 1], q.y, a1[c + 1]); }
        }
        double z0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), z1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
        const unsigned h0 = (unsigned)__double2hiint(z0) & 0x7fffffffu, h1 = (unsigned)__double2hiint(z1) & 0x7fffffffu;
#if TVF_PI_STICKY
        const unsigned hmax = half_max2_u32(max(h0, h1), lane & 16);
        double pv = shfl_d(pivA ? z0 : z1, src);
        const bool stale = ((unsigned)__double2hiint(pv) & 0x7fffffffu) + 0x00100000u < hmax;     // uniform within a half
        if (__any_sync(FULL, stale)) {
            const unsigned bal0 = __ballot_sync(FULL, h0 == hmax) & hmask;
            const unsigned bal1 = __ballot_sync(FULL, h1 == hmax) & hmask;
            if (stale) { pivA = (bal0 != 0u); src = __ffs(bal0 ? bal0 : bal1) - 1; }
            pv = shfl_d(pivA ? z0 : z1, src);
        }
#else
        const unsigned hmax = half_max_u32(max(h0, h1));
        const unsigned bal0 = __ballot_sync(FULL, h0 == hmax) & hmask;
        const unsigned bal1 = __ballot_sync(FULL, h1 == hmax) & hmask;
        const int src = __ffs(bal0 ? bal0 : bal1) - 1;
        const double pv = shfl_d(bal0 ? z0 : z1, src);
#endif
        const double ip = fast_rcp(pv);
        z0 *= ip; z1 *= ip;
        const bool moving = !done && ((fabs(fabs(z0) - fabs(x0)) > EIG_TOL) || (fabs(fabs(z1) - fabs(x1)) > EIG_TOL));
        if (!done) { x0 = z0; x1 = z1; }
        const unsigned bal = __ballot_sync(FULL, moving);
        done = done || ((bal & hmask) == 0u);
        if (bal == 0u) break;
    }
    const double nn = rsqrt_(half_sum(x0 * x0 + x1 * x1));
    *x0_out = x0 * nn; *x1_out = x1 * nn;
    *converged = done;
}

// ---------------------------------------------------------------------------------------------------------------
// One problem per warp, the matrix split over the lanes by ROWS AND COLUMNS: lane 16 h + r (r < 14) holds rows r and r + 14,
// columns [14 h, 14 h + 14) -- 28 doubles, as many as one full row.  The solver is bound by shared-memory delivery (a
// broadcast LDS.128 costs ~2 SM cycles, profiles/r02_microbench_smem_latency.txt): with one row per lane a sweep reads the
// whole 27-entry pivot row per lane (14 LDS.128); here a lane needs only the 14 entries of its column half (7 LDS.128, two
// addresses per instruction, same cost), its two multipliers (2 LDS.64) and the pivot (1 LDS.64): ~21 instead of ~29 SM
// cycles per sweep for the same 27-28 DFMAs, and the power iteration's mat-vec reads 7 LDS.128 and combines the two column
// halves with one shuffle pair.  The pivot column is published RAW (it is the pivot row by symmetry) and serves as row,
// multipliers and pivot at once; multipliers are scaled by 1/d per lane.  g0/g1: rows r / r + 14, PRESCALED (unit trace,
// shifted); entries of rows or columns >= N must be zero.  sbuf: 64 doubles.  Returns the components of rows r and r + 14.
template <int N>
__device__ __forceinline__ void smallest_eigvec_spd_cs(double (&g0)[14], double (&g1)[14], const int lane, double* sbuf,
                                                       double* x0_out, double* x1_out, bool* converged) {
    static_assert(N <= 28 && N > 14, "two rows x fourteen columns per lane");
    const int h = lane >> 4, r = lane & 15;
    const double floor_piv = 1.0e-3 * (1.0e-13 / N);
    const bool act = r < 14;
    const int rr = act ? r : 13;                       // lanes 14, 15 of each half idle (they mirror lane 13's loads)
    __syncwarp();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int hk = k / 14, kc = k % 14;            // column k: half hk, local column kc; pivot row k: lane kc, array hk
        double* vec = sbuf + (k & 1) * 32;
        if (h == hk && act) { vec[r] = g0[kc]; vec[14 + r] = g1[kc]; }
        __syncwarp();
        const double d = vec[k];
        const double c0 = vec[rr], c1 = vec[14 + rr];
        const double2* v2 = reinterpret_cast<const double2*>(vec + 14 * h);
        double2 q[7];
#pragma unroll
        for (int m = 0; m < 7; ++m) q[m] = v2[m];
        const double piv = fast_rcp(pivot_floor(d, floor_piv));
        const bool is_piv0 = (hk == 0) && (r == kc), is_piv1 = (hk == 1) && (r == kc);
        const double cs0 = (c0 - (is_piv0 ? 1.0 : 0.0)) * piv;
        const double cs1 = (c1 - (is_piv1 ? 1.0 : 0.0)) * piv;
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            g0[2 * m] = fma(-cs0, q[m].x, g0[2 * m]); g0[2 * m + 1] = fma(-cs0, q[m].y, g0[2 * m + 1]);
            g1[2 * m] = fma(-cs1, q[m].x, g1[2 * m]); g1[2 * m + 1] = fma(-cs1, q[m].y, g1[2 * m + 1]);
        }
        if (h == hk) {                                 // column k itself: -1/d on the pivot row, (column entry)/d elsewhere
            g0[kc] = is_piv0 ? -piv : c0 * piv;
            g1[kc] = is_piv1 ? -piv : c1 * piv;
        }
    }
    __syncwarp();
    // g holds -(G + delta I)^-1.  Pivot-normalised power iteration; both halves carry the same iterate.
    double x0 = act ? 1.0 : 0.0, x1 = (act && r + 14 < N) ? 1.0 : 0.0;
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < EIG_MAX_ITER; ++it) {
        double* vec = sbuf + (it & 1) * 32;
        if (h == 0 && act) { vec[r] = x0; vec[14 + r] = x1; }
        __syncwarp();
        const double2* v2 = reinterpret_cast<const double2*>(vec + 14 * h);
        double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const double2 qq = v2[m];
            a0[0] = fma(g0[2 * m], qq.x, a0[0]); a0[1] = fma(g0[2 * m + 1], qq.y, a0[1]);
            a1[0] = fma(g1[2 * m], qq.x, a1[0]); a1[1] = fma(g1[2 * m + 1], qq.y, a1[1]);
        }
        double z0 = a0[0] + a0[1], z1 = a1[0] + a1[1];
        z0 += __shfl_xor_sync(FULL, z0, 16); z1 += __shfl_xor_sync(FULL, z1, 16);       // the other column half
        const unsigned h0 = (unsigned)__double2hiint(z0) & 0x7fffffffu, h1 = (unsigned)__double2hiint(z1) & 0x7fffffffu;
        const unsigned hmax = __reduce_max_sync(FULL, max(h0, h1));
        const unsigned bal0 = __ballot_sync(FULL, h0 == hmax), bal1 = __ballot_sync(FULL, h1 == hmax);
        const int src = __ffs(bal0 ? bal0 : bal1) - 1;
        const double pv = shfl_d(bal0 ? z0 : z1, src);
        const double ip = fast_rcp(pv);
        z0 *= ip; z1 *= ip;
        const bool moving = (fabs(fabs(z0) - fabs(x0)) > EIG_TOL) || (fabs(fabs(z1) - fabs(x1)) > EIG_TOL);
        x0 = z0; x1 = z1;
        if (!__any_sync(FULL, moving)) { ok = true; break; }
    }
    const double nn = rsqrt_(half_sum(x0 * x0 + x1 * x1));
    *x0_out = x0 * nn; *x1_out = x1 * nn;
    *converged = ok;
}

}  // namespace tvf
