// tvf_core_kernels.cu -- warp-per-problem kernels for the two model estimators:
//   tft_core_kernel : Normalize2Ddata x3 (optional) -> 96 Kronecker moments ->
//                     27x27 Gram -> null vector -> epipoles -> 15-dim constrained
//                     re-solve -> (optional) undo normalisation
//                     (Normalize2Ddata.m:33-39, linearTFT.m:36-91,
//                      LinearTFTPoseEstimation.m:45-53)
//   f_core_kernel   : same idea for linearF.m:32-62 (36 moments, 9x9 Gram,
//                     rank-2 projection) and LinearFPoseEstimation.m:46-56
// One warp owns one triplet problem; the Gram never exists as a 4n x 27 design
// matrix: G = sum_i (p1 p1') (x) (S3 S3') (x) (S2 S2') is assembled from 96
// per-problem moments.
#include "tvf_kernels.h"
#include "tvf_warp.cuh"
#include "tvf_pose.cuh"

namespace tvf {

constexpr int CORE_WARPS = 8;
constexpr int FEAT_STRIDE = 15;   // 14 features + 1 pad: conflict-free 64-bit lane-strided stores

// per-warp shared scratch (doubles)
struct __align__(16) WarpScratch {
    double feat[32 * FEAT_STRIDE];   // per-point features; reused as W (27 x 15) in the constrained step
    double mom[98];                  // 96 moments + zero sentinel
    double T[28];                    // current tensor / F vector
    double vs[18];                   // slice null vectors
    double epi[6];                   // e21, e31
    double tp[16];
    double Nm[27];                   // N1, inv(N2), inv(N3)   (3x3 column-major each)
};

// index tables: which moment feeds G(r,c); 96 = structural zero
__device__ __constant__ unsigned char c_sym6[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
__device__ __constant__ signed char c_m4[9] = {0, -1, 1, -1, 0, 2, 1, 2, 3};

// v[i] for a runtime i without forcing the array into local memory
__device__ __forceinline__ double sel3(const double* v, int i) { return (i == 0) ? v[0] : ((i == 1) ? v[1] : v[2]); }

__device__ __forceinline__ void load_point(const CoreInput& in, long long prob, int i, double* p) {
    if (in.packed) {
        const double2* q = reinterpret_cast<const double2*>(in.p1 + (prob * in.n + i) * 6);
        const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y; p[4] = c.x; p[5] = c.y;
    } else {
        const long long o = (prob * in.n + i) * in.rows;
        const double* ps[3] = {in.p1, in.p2, in.p3};
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            if (ps[v] == nullptr) { p[2 * v] = 0.0; p[2 * v + 1] = 0.0; continue; }
            double x = ps[v][o], y = ps[v][o + 1];
            if (in.rows == 3) { const double w = ps[v][o + 2]; x /= w; y /= w; }   // linearTFT.m:39-43
            p[2 * v] = x; p[2 * v + 1] = y;
        }
    }
}

// Normalize2Ddata.m:34-37 for the three views at once, on points already mapped by
// x -> s0*x + t0 (identity when the outer map is absent).  Returns per view (sx, tx, ty)
// with new = s*x + t.
__device__ __forceinline__ void view_stats(const CoreInput& in, long long prob, int lane,
                                           const double* s0, const double* t0, double* s, double* t) {
    const int n = in.n;
    double sum[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        double p[6];
        load_point(in, prob, i, p);
#pragma unroll
        for (int q = 0; q < 6; ++q) sum[q] += s0[q >> 1] * p[q] + t0[q];
    }
    const double invn = 1.0 / (double)n;
    double c[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = warp_sum(sum[q]) * invn;
    double d[3] = {0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        double p[6];
        load_point(in, prob, i, p);
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            const double dx = (s0[v] * p[2 * v] + t0[2 * v]) - c[2 * v];
            const double dy = (s0[v] * p[2 * v + 1] + t0[2 * v + 1]) - c[2 * v + 1];
            d[v] += sqrt(dx * dx + dy * dy);
        }
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const double norm0 = warp_sum(d[v]) * invn;
        s[v] = 1.4142135623730951 / norm0;
        t[2 * v] = -s[v] * c[2 * v];
        t[2 * v + 1] = -s[v] * c[2 * v + 1];
    }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CORE_WARPS * 32, 2)
tft_core_kernel(CoreInput in, double* __restrict__ Tout, double* __restrict__ P2out,
                double* __restrict__ P3out, int* __restrict__ status) {
    __shared__ WarpScratch scratch[CORE_WARPS];
    __shared__ unsigned char gidx[32 * 27];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // G(r,c) -> moment index table (shared by all warps of the CTA)
    for (int e = threadIdx.x; e < 32 * 27; e += blockDim.x) {
        const int r = e / 27, c = e % 27;
        int idx = 96;
        if (r < 27) {
            const int j = r % 3, k = (r / 3) % 3, i = r / 9;
            const int j2 = c % 3, k2 = (c / 3) % 3, i2 = c / 9;
            const int g = c_m4[j * 3 + j2], b = c_m4[k * 3 + k2];
            if (g >= 0 && b >= 0) idx = c_sym6[i * 3 + i2] * 16 + b * 4 + g;
        }
        gidx[e] = (unsigned char)idx;
    }
    __syncthreads();

    WarpScratch& ws = scratch[warp];
    const int jr = lane % 3, kr = (lane / 3) % 3, ir = lane / 9;   // tensor indices of this lane's row (lane<27)

    for (long long prob = (long long)blockIdx.x * CORE_WARPS + warp; prob < in.B;
         prob += (long long)gridDim.x * CORE_WARPS) {
        int st = 0;
        // ---- normalisation (LinearTFTPoseEstimation.m:45-47) --------------------------
        double s[3] = {1.0, 1.0, 1.0}, t[6] = {0, 0, 0, 0, 0, 0};
        if (in.normalize) {
            const double s0[3] = {1.0, 1.0, 1.0}, t0[6] = {0, 0, 0, 0, 0, 0};
            view_stats(in, prob, lane, s0, t0, s, t);
        }
        // ---- 96 moments ----------------------------------------------------------------
        const int beta = (lane >> 2) & 3, gamma = lane & 3, alpha0 = lane >> 4;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
        for (int base = 0; base < in.n; base += 32) {
            const int cnt = min(32, in.n - base);
            __syncwarp();
            if (lane < cnt) {
                double p[6];
                load_point(in, prob, base + lane, p);
                const double x1 = s[0] * p[0] + t[0], y1 = s[0] * p[1] + t[1];
                const double x2 = s[1] * p[2] + t[2], y2 = s[1] * p[3] + t[3];
                const double x3 = s[2] * p[4] + t[4], y3 = s[2] * p[5] + t[5];
                double* f = ws.feat + lane * FEAT_STRIDE;
                f[0] = x1 * x1; f[1] = x1 * y1; f[2] = x1; f[3] = y1 * y1; f[4] = y1; f[5] = 1.0;
                f[6] = 1.0; f[7] = -x3; f[8] = -y3; f[9] = x3 * x3 + y3 * y3;
                f[10] = 1.0; f[11] = -x2; f[12] = -y2; f[13] = x2 * x2 + y2 * y2;
            }
            __syncwarp();
            for (int p = 0; p < cnt; ++p) {
                const double* f = ws.feat + p * FEAT_STRIDE;
                const double bc = f[6 + beta] * f[10 + gamma];
                acc0 = fma(f[alpha0], bc, acc0);
                acc1 = fma(f[alpha0 + 2], bc, acc1);
                acc2 = fma(f[alpha0 + 4], bc, acc2);
            }
        }
        __syncwarp();
        ws.mom[lane] = acc0; ws.mom[lane + 32] = acc1; ws.mom[lane + 64] = acc2;
        if (lane == 0) { ws.mom[96] = 0.0; ws.mom[97] = 0.0; }
        __syncwarp();

        // ---- stage 1: null vector of the 27x27 Gram (linearTFT.m:64-67) ------------------
        double g[27];
#pragma unroll
        for (int c = 0; c < 27; ++c) g[c] = ws.mom[gidx[lane * 27 + c]];
        bool conv;
        double tl = smallest_eigvec_spd<27>(g, lane, &conv);
        if (!conv) st |= ST_EIG_NOCONV;
        if (lane < 27) ws.T[lane] = tl;
        __syncwarp();

        // ---- epipoles (linearTFT.m:71-79): six slice problems on six lanes ---------------
        {
            const int l6 = lane % 6, sl = l6 % 3, tr = l6 / 3;
            double M[9], v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) M[r + 3 * c] = tr ? ws.T[9 * sl + c + 3 * r] : ws.T[9 * sl + r + 3 * c];
            null3(M, v);
            if (lane < 6) { ws.vs[3 * lane] = v[0]; ws.vs[3 * lane + 1] = v[1]; ws.vs[3 * lane + 2] = v[2]; }
            __syncwarp();
            const int which = lane & 1;           // 0: e31 from slices, 1: e21 from transposed slices
            double e[3];
            epipole_from_nulls(ws.vs + 9 * which, ws.vs + 9 * which + 3, ws.vs + 9 * which + 6, e);
            if (lane < 2) { ws.epi[3 * (1 - which)] = e[0]; ws.epi[3 * (1 - which) + 1] = e[1]; ws.epi[3 * (1 - which) + 2] = e[2]; }
            __syncwarp();
        }
        double e21[3] = {ws.epi[0], ws.epi[1], ws.epi[2]};
        double e31[3] = {ws.epi[3], ws.epi[4], ws.epi[5]};
        double u1[3], u2[3], v1[3], v2[3];
        onb3(e21, u1, u2);
        onb3(e31, v1, v2);

        // ---- stage 2: constrained re-solve in range(E) (linearTFT.m:82-85) ---------------
        // basis per slice: B0=e21 e31', B1=e21 v1', B2=e21 v2', B3=u1 e31', B4=u2 e31'
        {
            double Ze[9], Zu1[9], Zu2[9];   // Z_p[k'+3i'] = sum_j' G(r,(j',k',i')) p[j']
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const double g0 = ws.mom[gidx[lane * 27 + 3 * q]];
                const double g1 = ws.mom[gidx[lane * 27 + 3 * q + 1]];
                const double g2 = ws.mom[gidx[lane * 27 + 3 * q + 2]];
                Ze[q] = g0 * e21[0] + g1 * e21[1] + g2 * e21[2];
                Zu1[q] = g0 * u1[0] + g1 * u1[1] + g2 * u1[2];
                Zu2[q] = g0 * u2[0] + g1 * u2[1] + g2 * u2[2];
            }
            __syncwarp();   // feat no longer needed -> reuse as W
            if (lane < 27) {
                double* W = ws.feat + lane * FEAT_STRIDE;
#pragma unroll
                for (int i2 = 0; i2 < 3; ++i2) {
                    const double* ze = Ze + 3 * i2; const double* zu1 = Zu1 + 3 * i2; const double* zu2 = Zu2 + 3 * i2;
                    W[5 * i2 + 0] = ze[0] * e31[0] + ze[1] * e31[1] + ze[2] * e31[2];
                    W[5 * i2 + 1] = ze[0] * v1[0] + ze[1] * v1[1] + ze[2] * v1[2];
                    W[5 * i2 + 2] = ze[0] * v2[0] + ze[1] * v2[1] + ze[2] * v2[2];
                    W[5 * i2 + 3] = zu1[0] * e31[0] + zu1[1] * e31[1] + zu1[2] * e31[2];
                    W[5 * i2 + 4] = zu2[0] * e31[0] + zu2[1] * e31[1] + zu2[2] * e31[2];
                }
            }
            __syncwarp();
        }
        double tl2;
        {
            const int ia = (lane < 15) ? lane / 5 : 0, aa = (lane < 15) ? lane % 5 : 0;
            double pa[3], qa[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                pa[q] = (aa < 3) ? e21[q] : ((aa == 3) ? u1[q] : u2[q]);
                qa[q] = (aa == 1) ? v1[q] : ((aa == 2) ? v2[q] : e31[q]);
            }
            double g15[15];
#pragma unroll
            for (int c = 0; c < 15; ++c) g15[c] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double coef = pa[j] * qa[k];
                    const double* W = ws.feat + (9 * ia + 3 * k + j) * FEAT_STRIDE;
#pragma unroll
                    for (int c = 0; c < 15; ++c) g15[c] = fma(coef, W[c], g15[c]);
                }
            if (lane >= 15) {
#pragma unroll
                for (int c = 0; c < 15; ++c) g15[c] = 0.0;
            }
            bool conv2;
            const double tpl = smallest_eigvec_spd<15>(g15, lane, &conv2);
            if (!conv2) st |= ST_EIG_NOCONV;
            if (lane < 15) ws.tp[lane] = tpl;
            __syncwarp();
            // t = Up*tp  (linearTFT.m:85)
            double acc = 0.0;
            if (lane < 27) {
                const double* tp = ws.tp + 5 * ir;
                const double pe = sel3(e21, jr), pu1 = sel3(u1, jr), pu2 = sel3(u2, jr);
                const double qe = sel3(e31, kr), qv1 = sel3(v1, kr), qv2 = sel3(v2, kr);
                acc = pe * (qe * tp[0] + qv1 * tp[1] + qv2 * tp[2]) + qe * (pu1 * tp[3] + pu2 * tp[4]);
            }
            tl2 = acc * rsqrt(warp_sum(acc * acc));
        }
        __syncwarp();
        if (lane < 27) ws.T[lane] = tl2;
        __syncwarp();

        // ---- P2, P3 (linearTFT.m:86-90): a = pinv(E) t in closed form --------------------
        if (P2out != nullptr && P3out != nullptr) {
            if (lane < 3) {
                const double* Ti = ws.T + 9 * lane;
                double a[3], b[3];
                mat3_vec(Ti, e31, a);
                mat3_tvec(Ti, e21, b);
                const double tau = e21[0] * a[0] + e21[1] * a[1] + e21[2] * a[2];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    P2out[prob * 12 + 3 * lane + q] = a[q] - 0.5 * tau * e21[q];
                    P3out[prob * 12 + 3 * lane + q] = 0.5 * tau * e31[q] - b[q];
                }
            } else if (lane == 3) {
#pragma unroll
                for (int q = 0; q < 3; ++q) { P2out[prob * 12 + 9 + q] = e21[q]; P3out[prob * 12 + 9 + q] = e31[q]; }
            }
        }

        // ---- undo the normalisation (LinearTFTPoseEstimation.m:53) ----------------------
        double tout = tl2;
        if (in.normalize) {
            // N_v = [s 0 tx; 0 s ty; 0 0 1]; all lanes build the same matrices
            double N1[9] = {s[0], 0, 0, 0, s[0], 0, t[0], t[1], 1.0};
            double N2[9] = {s[1], 0, 0, 0, s[1], 0, t[2], t[3], 1.0};
            double N3[9] = {s[2], 0, 0, 0, s[2], 0, t[4], t[5], 1.0};
            double N2i[9], N3i[9];
            inv3(N2, N2i); inv3(N3, N3i);
            // dynamic register indexing is avoided by staging through shared memory
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) { ws.Nm[q] = N1[q]; ws.Nm[9 + q] = N2i[q]; ws.Nm[18 + q] = N3i[q]; }
            }
            __syncwarp();
            double acc = 0.0;
            if (lane < 27) {
                // T_new(j,k,i) = sum_{a,b} N2i(j,a) * (sum_r N1(r,i) T(a,b,r)) * N3i(k,b)   (transform_TFT.m:43-46)
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double sab = ws.Nm[3 * ir] * ws.T[a + 3 * b] + ws.Nm[1 + 3 * ir] * ws.T[9 + a + 3 * b] +
                                           ws.Nm[2 + 3 * ir] * ws.T[18 + a + 3 * b];
                        acc = fma(ws.Nm[9 + jr + 3 * a] * ws.Nm[18 + kr + 3 * b], sab, acc);
                    }
            }
            tout = acc * rsqrt(warp_sum(acc * acc));                                         // :49
        }
        if (lane < 27) Tout[prob * 27 + lane] = tout;
        if (status != nullptr && lane == 0) status[prob] = st;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// linearF.m:45-62 for one view pair inside the warp.  (s0,t0): outer map applied to the raw
// points of the two views (identity for a direct linearF call).  Writes the 3x3 F
// (column-major) into Fout (all lanes hold it).
__device__ __forceinline__ int linear_f_pair(const CoreInput& in, long long prob, int lane, WarpScratch& ws,
                                             int va, int vb, const double* s0, const double* t0,
                                             const double* si, const double* ti, double* F) {
    // inner map composed with outer: x' = si*(s0*x + t0) + ti
    const double sa = si[va] * s0[va], sb = si[vb] * s0[vb];
    const double tax = si[va] * t0[2 * va] + ti[2 * va], tay = si[va] * t0[2 * va + 1] + ti[2 * va + 1];
    const double tbx = si[vb] * t0[2 * vb] + ti[2 * vb], tby = si[vb] * t0[2 * vb + 1] + ti[2 * vb + 1];
    // 36 moments: lane l -> (alpha,beta) = (l/6, l%6); lanes 0..3 also take moment 32+l
    const int m0 = lane, m1 = 32 + lane;
    const int a0 = m0 / 6, b0 = m0 % 6, a1 = (m1 < 36) ? m1 / 6 : 0, b1 = (m1 < 36) ? m1 % 6 : 0;
    double acc0 = 0.0, acc1 = 0.0;
    for (int base = 0; base < in.n; base += 32) {
        const int cnt = min(32, in.n - base);
        __syncwarp();
        if (lane < cnt) {
            double p[6];
            load_point(in, prob, base + lane, p);
            const double x1 = sa * p[2 * va] + tax, y1 = sa * p[2 * va + 1] + tay;
            const double x2 = sb * p[2 * vb] + tbx, y2 = sb * p[2 * vb + 1] + tby;
            double* f = ws.feat + lane * FEAT_STRIDE;
            f[0] = x1 * x1; f[1] = x1 * y1; f[2] = x1; f[3] = y1 * y1; f[4] = y1; f[5] = 1.0;
            f[6] = x2 * x2; f[7] = x2 * y2; f[8] = x2; f[9] = y2 * y2; f[10] = y2; f[11] = 1.0;
        }
        __syncwarp();
        for (int p = 0; p < cnt; ++p) {
            const double* f = ws.feat + p * FEAT_STRIDE;
            acc0 = fma(f[a0], f[6 + b0], acc0);
            acc1 = fma(f[a1], f[6 + b1], acc1);
        }
    }
    __syncwarp();
    ws.mom[lane] = acc0;
    if (lane < 4) ws.mom[32 + lane] = acc1;
    __syncwarp();
    // G9(r,c): r = 3*a + b with A row [x1x2, x1y2, x1, y1x2, y1y2, y1, x2, y2, 1]   (linearF.m:51-52)
    double g[9];
    {
        const int ar = (lane < 9) ? lane / 3 : 0, br = (lane < 9) ? lane % 3 : 0;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const int ac = c / 3, bc = c % 3;
            const double v = ws.mom[c_sym6[ar * 3 + ac] * 6 + c_sym6[br * 3 + bc]];
            g[c] = (lane < 9) ? v : 0.0;
        }
    }
    bool conv;
    const double fl = smallest_eigvec_spd<9>(g, lane, &conv);
    __syncwarp();
    if (lane < 9) ws.T[lane] = fl;
    __syncwarp();
    // F = reshape(V(:,9),3,3); F = Normal2.'*F*Normal1 (inner maps only); rank-2 projection (:55-62)
    double Fv[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) Fv[q] = ws.T[q];
    const double Na[9] = {si[va], 0, 0, 0, si[va], 0, ti[2 * va], ti[2 * va + 1], 1.0};
    const double Nb[9] = {si[vb], 0, 0, 0, si[vb], 0, ti[2 * vb], ti[2 * vb + 1], 1.0};
    double tmp[9], Fu[9];
    mat3_mul_tn(Nb, Fv, tmp);
    mat3_mul(tmp, Na, Fu);
    double U[9], sv[3], V[9];
    svd3_full(Fu, U, sv, V);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) F[r + 3 * c] = sv[0] * U[r] * V[c] + sv[1] * U[3 + r] * V[3 + c];
    return conv ? 0 : ST_EIG_NOCONV;
}

// mode 0: direct linearF(p1,p2) -> Fout[9*prob];  mode 1: pose path, F21 -> Fout[18*prob], F31 -> +9
__global__ void __launch_bounds__(CORE_WARPS * 32, 2)
f_core_kernel(CoreInput in, double* __restrict__ Fout, int* __restrict__ status) {
    __shared__ WarpScratch scratch[CORE_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    for (long long prob = (long long)blockIdx.x * CORE_WARPS + warp; prob < in.B;
         prob += (long long)gridDim.x * CORE_WARPS) {
        int st = 0;
        const double one3[3] = {1.0, 1.0, 1.0}, zero6[6] = {0, 0, 0, 0, 0, 0};
        double s0[3] = {1.0, 1.0, 1.0}, t0[6] = {0, 0, 0, 0, 0, 0};
        if (in.normalize) view_stats(in, prob, lane, one3, zero6, s0, t0);   // LinearFPoseEstimation.m:46-48
        double si[3], ti[6];
        view_stats(in, prob, lane, s0, t0, si, ti);                          // linearF.m:45-46 (re-normalisation)
        const int npairs = in.normalize ? 2 : 1;
        for (int pr = 0; pr < npairs; ++pr) {
            const int vb = 1 + pr;
            double F[9];
            st |= linear_f_pair(in, prob, lane, ws, 0, vb, s0, t0, si, ti, F);
            if (in.normalize) {                                              // LinearFPoseEstimation.m:55-56
                const double Na[9] = {s0[0], 0, 0, 0, s0[0], 0, t0[0], t0[1], 1.0};
                const double Nb[9] = {s0[vb], 0, 0, 0, s0[vb], 0, t0[2 * vb], t0[2 * vb + 1], 1.0};
                double tmp[9], Fo[9];
                mat3_mul_tn(Nb, F, tmp);
                mat3_mul(tmp, Na, Fo);
#pragma unroll
                for (int q = 0; q < 9; ++q) F[q] = Fo[q];
            }
            // lane q writes F[q] without dynamic register indexing
            double mine = 0.0;
#pragma unroll
            for (int q = 0; q < 9; ++q) mine = (lane == q) ? F[q] : mine;
            if (lane < 9) Fout[prob * (9 * npairs) + 9 * pr + lane] = mine;
        }
        if (status != nullptr && lane == 0) status[prob] = st;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
void launch_tft_core(const CoreInput& in, double* T, double* P2, double* P3, int* status, int sm_count,
                     cudaStream_t stream) {
    if (in.B <= 0) return;
    long long blocks = (in.B + CORE_WARPS - 1) / CORE_WARPS;
    const long long cap = (long long)sm_count * 2 * 8;
    if (blocks > cap) blocks = cap;
    tft_core_kernel<<<(unsigned)blocks, CORE_WARPS * 32, 0, stream>>>(in, T, P2, P3, status);
}

void launch_f_core(const CoreInput& in, double* F, int* status, int sm_count, cudaStream_t stream) {
    if (in.B <= 0) return;
    long long blocks = (in.B + CORE_WARPS - 1) / CORE_WARPS;
    const long long cap = (long long)sm_count * 2 * 8;
    if (blocks > cap) blocks = cap;
    f_core_kernel<<<(unsigned)blocks, CORE_WARPS * 32, 0, stream>>>(in, F, status);
}

}  // namespace tvf
