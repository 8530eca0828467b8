// tvf_gh_kernels.cu -- Gauss-Helmert refinement of the fundamental matrix (first step):
// F_methods/optimF.m:50-76 with Optimization/Gauss_Helmert.m:38-83 specialised to its constraint function
// constraintsGH_F (optimF.m:81-109).  One warp owns one (problem, view pair); lanes own points.
//
// What the specialisation uses (and nothing else):
//  * P = eye(4N) (optimF.m:65), so pinv(P) = inv(P) = I.
//  * B is block diagonal with one 1x4 block b_i per point (optimF.m:103-104), hence W = B*B' is DIAGONAL:
//    W_ii = |b_i|^2, and Gauss_Helmert.m:57  pinv(W + 1e-12 I) + 1e-12 I  is the per-point weight
//    w_i = 1/(|b_i|^2 + 1e-12) + 1e-12  (entries not above MATLAB pinv's tolerance N*eps(max_i W_ii) become 0).
//    The reference forms the dense N x 4N / N x N / 4N x 4N matrices; none of them exists here.
//  * y is empty (optimF.m:64), so M = [A'WA C'; C 0] is 11 x 11 (9 parameters + det(F) = 0 and |F|^2 = 1).
//    Gauss_Helmert.m:67 applies pinv(M + 1e-12 I); M + 1e-12 I is non-singular (the constraint gradients span
//    the null direction of A'WA), so this is a linear solve: Gauss-Jordan with partial pivoting in shared memory.
//  * v = -B'(W(A dt - w)) is per point v_i = -b_i * w_i * (a_i.dt - w~_i)  (:69).
// The control flow (tolerances, the two break tests, no update on the breaking iteration, it_max = 400,
// NaN/Inf guards) follows Gauss_Helmert.m:49-81 statement by statement; `iter` is the loop counter at exit.
#include "tvf_kernels.h"
#include "tvf_warp.cuh"
#include "tvf_pose.cuh"

namespace tvf {

constexpr int GH_WARPS = 4;
constexpr int GH_IT_MAX = 400;            // Gauss_Helmert.m:38
constexpr double GH_TOL = 1e-6;           // :39
constexpr int GH_FEAT = 19;               // per-point record: w*a (9) | a (9) | w~ ; odd stride -> conflict-free
constexpr int FW_STATS_GH = 18;           // CORE_WS_F record: raw f (2 x 9) | outer stats s0[3], t0[6] | inner stats

// entry e of the 54 sums a warp accumulates per iteration: 45 upper-triangular entries of A'WA, then A'W w~
__device__ __forceinline__ void gh_entry(int e, int& r, int& c) {
    if (e >= 45) { r = e - 45; c = 9; return; }
    int rr = 0, base = 0;
    while (e >= base + (9 - rr)) { base += 9 - rr; ++rr; }
    r = rr; c = rr + (e - base);
}

__device__ __forceinline__ double eps_of(double x) {      // MATLAB eps(x) for a positive normal x
    const long long b = __double_as_longlong(x) & 0x7ff0000000000000LL;
    return __longlong_as_double(b - (52LL << 52));
}

struct GhPoint { double a[9], b[4], f; };

// constraintsGH_F (optimF.m:99-104) for one point: x = (x1x, x1y, x2x, x2y), F column-major
__device__ __forceinline__ void gh_point(const double* x, const double* F, GhPoint& p) {
    const double x1x = x[0], x1y = x[1], x2x = x[2], x2y = x[3];
    p.a[0] = x1x * x2x; p.a[1] = x1x * x2y; p.a[2] = x1x; p.a[3] = x1y * x2x; p.a[4] = x1y * x2y; p.a[5] = x1y;
    p.a[6] = x2x; p.a[7] = x2y; p.a[8] = 1.0;                                                   // :102
    double f = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) f = fma(p.a[k], F[k], f);                                       // x2.'*F*x1  (:101)
    p.f = f;
    p.b[0] = F[2] + F[0] * x2x + F[1] * x2y; p.b[1] = F[5] + F[3] * x2x + F[4] * x2y;           // :103
    p.b[2] = F[6] + F[0] * x1x + F[3] * x1y; p.b[3] = F[7] + F[1] * x1x + F[4] * x1y;           // :104
}

// dynamic shared memory per warp: xo[4n] | xa[4n] | xb[4n] | feat[32*19] | M[11*12] | dt[12]
__global__ void __launch_bounds__(GH_WARPS * 32)
optimf_gh_kernel(const double* __restrict__ corresp, int n, long long B, const double* __restrict__ ws,
                 double* __restrict__ Fio, int* __restrict__ iters, int* __restrict__ status) {
    extern __shared__ __align__(16) double gh_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = 12 * n + 32 * GH_FEAT + 11 * 12 + 12;
    double* xo = gh_smem + (size_t)warp * per_warp;
    double* xcur = xo + 4 * n;
    double* xnew = xcur + 4 * n;
    double* feat = xnew + 4 * n;
    double* M = feat + 32 * GH_FEAT;
    double* dts = M + 11 * 12;
    int r0, c0, r1, c1;
    gh_entry(lane, r0, c0);
    gh_entry(min(lane + 32, 53), r1, c1);
    const bool has1 = lane + 32 < 54;

    for (long long unit = (long long)blockIdx.x * GH_WARPS + warp; unit < 2 * B; unit += (long long)gridDim.x * GH_WARPS) {
        const long long prob = unit >> 1;
        const int pr = (int)(unit & 1), vb = 1 + pr;
        const double* rec = ws + prob * CORE_WS_F + FW_STATS_GH;      // s0[3], t0[6]: Normalize2Ddata of optimF.m:46-47
        const double sa = rec[0], sb = rec[vb];
        const double tax = rec[3], tay = rec[4], tbx = rec[3 + 2 * vb], tby = rec[4 + 2 * vb];
        double F[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) F[k] = Fio[prob * 18 + 9 * pr + k];      // linearF, unit Frobenius norm (:50)
        __syncwarp();
        // ---- optimF.m:53-60: cameras from F, triangulation, reprojected points = first estimate ----------
        double e[3];
        null3_t(F, e);                                               // U(:,3) of svd(F) (:53), up to sign
        double P1[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}, P2[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {                                // crossM(epi21)*F (:55)
            const double f0 = F[3 * c], f1 = F[1 + 3 * c], f2 = F[2 + 3 * c];
            P2[3 * c] = e[1] * f2 - e[2] * f1; P2[1 + 3 * c] = e[2] * f0 - e[0] * f2; P2[2 + 3 * c] = e[0] * f1 - e[1] * f0;
        }
        P2[9] = e[0]; P2[10] = e[1]; P2[11] = e[2];
        double obj_l = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double2* q = reinterpret_cast<const double2*>(corresp + (prob * n + i) * 6);
            const double2 pa = __ldg(q), pb = __ldg(q + vb);
            const double x1x = sa * pa.x + tax, x1y = sa * pa.y + tay, x2x = sb * pb.x + tbx, x2y = sb * pb.y + tby;
            xo[4 * i] = x1x; xo[4 * i + 1] = x1y; xo[4 * i + 2] = x2x; xo[4 * i + 3] = x2y;
            double rows[4][4];
            dlt_rows(P1, x1x, x1y, rows[0], rows[1]);
            dlt_rows(P2, x2x, x2y, rows[2], rows[3]);
            double X[4], h[3];
            dlt_null<4>(rows, X);                                    // triangulation3D (:56)
            const double e1x = X[0] / X[2], e1y = X[1] / X[2];       // P1*X (:59)
            cam_apply(P2, X, h);
            const double e2x = h[0] / h[2], e2y = h[1] / h[2];       // (:60)
            xcur[4 * i] = e1x; xcur[4 * i + 1] = e1y; xcur[4 * i + 2] = e2x; xcur[4 * i + 3] = e2y;
            const double d0 = e1x - x1x, d1 = e1y - x1y, d2 = e2x - x2x, d3 = e2y - x2y;
            obj_l += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        double objFunc = warp_sum(obj_l);                            // Gauss_Helmert.m:45-46
        __syncwarp();
        int it = 0;
        int st = 0;
        for (it = 1; it <= GH_IT_MAX; ++it) {                        // :49
            // ---- W = B*B' (diagonal), NaN/Inf guard (:52-55), pinv tolerance --------------------------------
            double dmax_l = 0.0; bool bad_l = false;
            for (int i = lane; i < n; i += 32) {
                GhPoint p; gh_point(xcur + 4 * i, F, p);
                const double d = p.b[0] * p.b[0] + p.b[1] * p.b[1] + p.b[2] * p.b[2] + p.b[3] * p.b[3];
                bad_l = bad_l || !(fabs(d) <= 1.79769313486231570e308);
                dmax_l = fmax(dmax_l, d);
            }
            if (__any_sync(FULL, bad_l)) break;
            const double dmax = warp_max(dmax_l) + 1e-12;
            const double ptol = (double)n * eps_of(dmax);            // pinv(W + 1e-12 I): tol = max(size)*eps(norm)
            // ---- A'WA and A'W w~ -----------------------------------------------------------------------------
            double acc0 = 0.0, acc1 = 0.0;
            for (int pbase = 0; pbase < n; pbase += 32) {
                const int cnt = min(32, n - pbase);
                __syncwarp();
                if (lane < cnt) {
                    const int i = pbase + lane;
                    GhPoint p; gh_point(xcur + 4 * i, F, p);
                    const double d = p.b[0] * p.b[0] + p.b[1] * p.b[1] + p.b[2] * p.b[2] + p.b[3] * p.b[3] + 1e-12;
                    const double wi = ((d > ptol) ? 1.0 / d : 0.0) + 1e-12;                      // :57
                    const double wt = -p.f - (p.b[0] * (xo[4 * i] - xcur[4 * i]) + p.b[1] * (xo[4 * i + 1] - xcur[4 * i + 1]) +
                                              p.b[2] * (xo[4 * i + 2] - xcur[4 * i + 2]) + p.b[3] * (xo[4 * i + 3] - xcur[4 * i + 3]));   // :58
                    double* fr = feat + lane * GH_FEAT;
#pragma unroll
                    for (int k = 0; k < 9; ++k) { fr[k] = wi * p.a[k]; fr[9 + k] = p.a[k]; }
                    fr[18] = wt;
                }
                __syncwarp();
                for (int p = 0; p < cnt; ++p) {
                    const double* fr = feat + p * GH_FEAT;
                    acc0 = fma(fr[r0], fr[9 + c0], acc0);
                    acc1 = fma(fr[r1], fr[9 + c1], acc1);
                }
            }
            // ---- M = [A'WA C'; C 0] + 1e-12 I, b = [A'W w~; -g]  (:59-62, 67) ------------------------------------
            __syncwarp();
            if (c0 < 9) { M[r0 * 12 + c0] = acc0 + ((r0 == c0) ? 1e-12 : 0.0); M[c0 * 12 + r0] = acc0 + ((r0 == c0) ? 1e-12 : 0.0); }
            else M[r0 * 12 + 11] = acc0;
            if (has1) {
                if (c1 < 9) { M[r1 * 12 + c1] = acc1 + ((r1 == c1) ? 1e-12 : 0.0); M[c1 * 12 + r1] = acc1 + ((r1 == c1) ? 1e-12 : 0.0); }
                else M[r1 * 12 + 11] = acc1;
            }
            double sumsq = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) sumsq += F[k] * F[k];
            const double Cd[9] = {F[4] * F[8] - F[5] * F[7], F[5] * F[6] - F[3] * F[8], F[3] * F[7] - F[4] * F[6],
                                  F[2] * F[7] - F[1] * F[8], F[0] * F[8] - F[2] * F[6], F[1] * F[6] - F[0] * F[7],
                                  F[1] * F[5] - F[2] * F[4], F[2] * F[3] - F[0] * F[5], F[0] * F[4] - F[1] * F[3]};    // optimF.m:90-92
            if (lane < 9) {
                double cd = 0.0, fk = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) { cd = (lane == k) ? Cd[k] : cd; fk = (lane == k) ? F[k] : fk; }
                M[9 * 12 + lane] = cd; M[lane * 12 + 9] = cd;
                M[10 * 12 + lane] = 2.0 * fk; M[lane * 12 + 10] = 2.0 * fk;                     // optimF.m:93
            }
            if (lane == 9) {
                M[9 * 12 + 9] = 1e-12; M[9 * 12 + 10] = 0.0; M[10 * 12 + 9] = 0.0; M[10 * 12 + 10] = 1e-12;
                M[9 * 12 + 11] = -det3(F);                                                     // -g (optimF.m:88)
                M[10 * 12 + 11] = -(sumsq - 1.0);
            }
            __syncwarp();
            bool badM = false;
            for (int q = lane; q < 11 * 12; q += 32) badM = badM || !(fabs(M[q]) <= 1.79769313486231570e308);
            if (__any_sync(FULL, badM)) break;                                                  // :63-65
            // ---- aux = (M + 1e-12 I) \ b: Gauss-Jordan, partial pivoting, lane r = row r ------------------------
            int sing = 0;
#pragma unroll 1
            for (int k = 0; k < 11; ++k) {
                double av = (lane >= k && lane < 11) ? fabs(M[lane * 12 + k]) : -1.0;
                int ai = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(FULL, av, o);
                    const int oi = __shfl_xor_sync(FULL, ai, o);
                    if (ov > av || (ov == av && oi < ai)) { av = ov; ai = oi; }
                }
                if (!(av > 0.0)) { sing = 1; break; }
                if (ai != k && lane < 12) { const double t = M[k * 12 + lane]; M[k * 12 + lane] = M[ai * 12 + lane]; M[ai * 12 + lane] = t; }
                __syncwarp();
                const double piv = M[k * 12 + k];
                if (lane < 11 && lane != k) {
                    const double fct = M[lane * 12 + k] / piv;
                    for (int c = k + 1; c < 12; ++c) M[lane * 12 + c] = fma(-fct, M[k * 12 + c], M[lane * 12 + c]);
                }
                __syncwarp();
            }
            if (sing) { st |= ST_EIG_NOCONV; break; }
            if (lane < 9) dts[lane] = M[lane * 12 + 11] / M[lane * 12 + lane];                  // dt (:68)
            __syncwarp();
            double dt[9], ndt2 = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) { dt[k] = dts[k]; ndt2 += dt[k] * dt[k]; }
            // ---- v = -B'(W(A dt - w~))  (:69) and the quantities of the two tests ------------------------------
            double vv_l = 0.0, res_l = 0.0;
            for (int i = lane; i < n; i += 32) {
                GhPoint p; gh_point(xcur + 4 * i, F, p);
                const double d = p.b[0] * p.b[0] + p.b[1] * p.b[1] + p.b[2] * p.b[2] + p.b[3] * p.b[3] + 1e-12;
                const double wi = ((d > ptol) ? 1.0 / d : 0.0) + 1e-12;
                const double wt = -p.f - (p.b[0] * (xo[4 * i] - xcur[4 * i]) + p.b[1] * (xo[4 * i + 1] - xcur[4 * i + 1]) +
                                          p.b[2] * (xo[4 * i + 2] - xcur[4 * i + 2]) + p.b[3] * (xo[4 * i + 3] - xcur[4 * i + 3]));
                double adt = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) adt = fma(p.a[k], dt[k], adt);
                const double cf = -(wi * (adt - wt));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double v = p.b[k] * cf;
                    vv_l += v * v;
                    const double rr = xcur[4 * i + k] - xo[4 * i + k] - v;
                    res_l += rr * rr;
                    xnew[4 * i + k] = xo[4 * i + k] + v;                                        // xi = x + v (:80), if accepted
                }
            }
            const double vv = warp_sum(vv_l), res = warp_sum(res_l);
            if (sqrt(ndt2) < GH_TOL && sqrt(res) < GH_TOL) break;                               // :71-73 (dy is empty)
            if (vv > objFunc) break;                                                            // :75-76 (factor = 1)
            objFunc = vv;                                                                       // :78
            __syncwarp();
            { double* t = xcur; xcur = xnew; xnew = t; }
#pragma unroll
            for (int k = 0; k < 9; ++k) F[k] += dt[k];                                          // ti = ti + dt (:80)
        }
        if (it > GH_IT_MAX) it = GH_IT_MAX;                                                     // loop ran out: iter = it_max
        // ---- optimF.m:69-76: undo the normalisation, rank-2 projection -----------------------------------------
        const double Na[9] = {sa, 0, 0, 0, sa, 0, tax, tay, 1.0};
        const double Nb[9] = {sb, 0, 0, 0, sb, 0, tbx, tby, 1.0};
        double tmp[9], Fu[9], U[9], sv[3], V[9];
        mat3_mul_tn(Nb, F, tmp);
        mat3_mul(tmp, Na, Fu);                                                                  // :72
        svd3_full(Fu, U, sv, V);                                                                // :75
        if (lane < 9) {
            const int r = lane % 3, c = lane / 3;
            double ur0 = 0, ur1 = 0, vc0 = 0, vc1 = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ur0 = (r == k) ? U[k] : ur0; ur1 = (r == k) ? U[3 + k] : ur1;
                vc0 = (c == k) ? V[k] : vc0; vc1 = (c == k) ? V[3 + k] : vc1;
            }
            Fio[prob * 18 + 9 * pr + lane] = sv[0] * ur0 * vc0 + sv[1] * ur1 * vc1;             // :76
        }
        if (lane == 0) {
            iters[2 * prob + pr] = it;                                                          // optimF.m:66 -> iter
            if (st != 0 && status != nullptr) atomicOr(&status[prob], st);
        }
        __syncwarp();
    }
}

// linearF result in the outer-normalised frame, scaled to unit Frobenius norm (optimF.m:50): the starting point of
// the refinement.  One thread per problem; reads the f_stage1 records, writes F0 for both pairs (18 per problem).
__global__ void __launch_bounds__(128)
optimf_init_kernel(const double* __restrict__ ws, long long B, double* __restrict__ Fout) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* rec = ws + b * CORE_WS_F;
    const double* si = rec + FW_STATS_GH + 9; const double* ti = si + 3;
    for (int pr = 0; pr < 2; ++pr) {
        const int vb = 1 + pr;
        double Fv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) Fv[q] = rec[9 * pr + q];                                    // reshape(V(:,9),3,3)
        const double Na[9] = {si[0], 0, 0, 0, si[0], 0, ti[0], ti[1], 1.0};
        const double Nb[9] = {si[vb], 0, 0, 0, si[vb], 0, ti[2 * vb], ti[2 * vb + 1], 1.0};
        double tmp[9], Fu[9], F[9], U[9], sv[3], V[9];
        mat3_mul_tn(Nb, Fv, tmp);
        mat3_mul(tmp, Na, Fu);                                                                  // linearF.m:58
        svd3_full(Fu, U, sv, V);                                                                // linearF.m:61
        double nrm = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                F[r + 3 * c] = sv[0] * U[r] * V[c] + sv[1] * U[3 + r] * V[3 + c];               // linearF.m:62
                nrm += F[r + 3 * c] * F[r + 3 * c];
            }
        const double inv = 1.0 / sqrt(nrm);                                                     // optimF.m:50
#pragma unroll
        for (int q = 0; q < 9; ++q) Fout[b * 18 + 9 * pr + q] = F[q] * inv;
    }
}

__global__ void sum_pairs_kernel(const int* __restrict__ in2, long long B, int* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) out[b] = in2[2 * b] + in2[2 * b + 1];                       // OptimFPoseEstimation.m:49
}

void launch_sum_pairs(const int* in2, long long B, int* out, cudaStream_t stream) {
    if (B > 0) sum_pairs_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(in2, B, out);
}

int optimf_max_n() {      // 4 warps x (12 n + 776) doubles of shared memory must fit 200 KB
    return (int)((200 * 1024 / (GH_WARPS * sizeof(double)) - (32 * GH_FEAT + 11 * 12 + 12)) / 12);
}

int launch_optimf_gh(const double* corresp, int n, long long B, const double* ws, double* Fio, int* iters, int* status,
                     int sm_count, cudaStream_t stream) {
    if (B <= 0) return 1;
    const size_t smem = (size_t)GH_WARPS * (12 * (size_t)n + 32 * GH_FEAT + 11 * 12 + 12) * sizeof(double);
    if (smem > 200 * 1024) return 0;
    // the opt-in is a per-device (per-context) property of the function: set it on every launch (cheap), never cached
    // in a process-wide static -- a handle on a second GPU would otherwise never get it
    if (cudaFuncSetAttribute(optimf_gh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    optimf_init_kernel<<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(ws, B, Fio);
    long long blocks = (2 * B + GH_WARPS - 1) / GH_WARPS;
    const long long cap = (long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    optimf_gh_kernel<<<(unsigned)blocks, GH_WARPS * 32, smem, stream>>>(corresp, n, B, ws, Fio, iters, status);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : 0;
}

}  // namespace tvf
