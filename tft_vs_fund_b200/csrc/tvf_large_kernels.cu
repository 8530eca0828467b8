// tvf_large_kernels.cu -- Gram formation for large-n triplets (BASELINE config 5: 10 000 correspondences
// per scene): the first half of linearTFT's stage 1 (Normalize2Ddata.m:33-39 x3 + the design-matrix Gram
// of linearTFT.m:45-64) as a bandwidth-oriented kernel.
//
// One thread-block CLUSTER of 8 CTAs owns a scene.  Each CTA pulls its contiguous slice of the scene
// (n/8 points x 48 B) from HBM into shared memory with ONE bulk asynchronous copy (TMA, cp.async.bulk ->
// SASS UBLKCP) that signals an mbarrier; three CTAs of different clusters share an SM, so one CTA's copy and
// cluster barriers overlap the arithmetic of the others.  The scene is read from HBM exactly once; the three passes the reference's normalisation
// forces (mean -> mean distance -> moments of the normalised points) run on the shared-memory copy, and
// only 6 + 3 + 96 partial sums per CTA cross the cluster through distributed shared memory (DSMEM),
// added in rank order (deterministic).
//
// Per point the arithmetic is ~170 FP64 lane-operations against 48 bytes: ~3.5 flop/B executed, right on
// the B200 FP64/HBM ridge (64 DFMA/clk/SM vs ~30 B/clk/SM), which is why this step is reported against both
// rooflines.  Tensor cores are deliberately not used: the Kronecker form needs 96 accumulations per point,
// a dense FP64 MMA on the 4x27 rows would need 5 832.
#include <cooperative_groups.h>

#include "tvf_kernels.h"
#include "tvf_warp.cuh"

namespace cg = cooperative_groups;

namespace tvf {

constexpr int LG_CLUSTER = 8;
constexpr int LG_THREADS = 128;
constexpr int LG_WARPS = LG_THREADS / 32;
constexpr int CW_MOM_L = 0, CW_STATS_L = 96;

// ---- mbarrier / bulk-copy PTX --------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double warp_sum_l(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// shared-memory layout (dynamic): [slice bytes] | Scratch
struct __align__(16) LargeScratch {
    unsigned long long bar;
    double wpart[LG_WARPS][32];     // per-warp partials (passes 1, 2) / transposed moment sums (pass 3)
    double slot1[2][8];             // this CTA's partial sums, double-buffered by scene parity: coordinates (6)
    double slot2[2][4];             //                                                          : distances (3)
    double slot3[2][96];            //                                                          : moments
    double bcast[12];               // cluster totals fetched by a few threads, read by all (centroids / scales)
};

// grid = num_clusters * 8 CTAs; cluster c handles scenes c, c + num_clusters, ...  Three CTAs (of different
// clusters) share an SM, so one CTA's bulk copy and cluster barriers overlap the arithmetic of the others.
__global__ void __cluster_dims__(LG_CLUSTER, 1, 1) __launch_bounds__(LG_THREADS, 3)
tft_moments_large_kernel(const double* __restrict__ corresp, int n, long long B, int slice_pts, int normalize,
                         double* __restrict__ ws) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const long long cid = blockIdx.x / LG_CLUSTER, ncl = gridDim.x / LG_CLUSTER;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t slice_bytes = (size_t)slice_pts * 48;
    const double* pts = reinterpret_cast<const double*>(smem_raw);
    LargeScratch& sc = *reinterpret_cast<LargeScratch*>(smem_raw + slice_bytes);

    const int p_lo = min(n, rank * slice_pts), p_hi = min(n, p_lo + slice_pts);
    const int npts = p_hi - p_lo;                       // this CTA's points of every scene
    const unsigned bytes = (unsigned)npts * 48u;
    const double inv_n = 1.0 / (double)n;

    if (tid == 0) {
        mbar_init(&sc.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase = 0u;
    int par = 0;
    for (long long scene = cid; scene < B; scene += ncl, par ^= 1) {
        // the previous scene's last cluster.sync guarantees every thread of this CTA is done with the buffer
        if (bytes > 0) {
            if (tid == 0) {
                mbar_expect_tx(&sc.bar, bytes);
                bulk_g2s(smem_raw, corresp + (scene * n + p_lo) * 6, bytes, &sc.bar);
            }
            mbar_wait(&sc.bar, phase); phase ^= 1u;
        }
        // ---- pass 1: centroids (Normalize2Ddata.m:34) ------------------------------------------------
        double s[3] = {1.0, 1.0, 1.0}, t[6] = {0, 0, 0, 0, 0, 0};
        if (normalize) {
            double sum[6] = {0, 0, 0, 0, 0, 0};
            for (int i = tid; i < npts; i += LG_THREADS) {
                const double2* q = reinterpret_cast<const double2*>(pts + 6 * i);
                const double2 a = q[0], b = q[1], c = q[2];
                sum[0] += a.x; sum[1] += a.y; sum[2] += b.x; sum[3] += b.y; sum[4] += c.x; sum[5] += c.y;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) { const double w = warp_sum_l(sum[k]); if (lane == 0) sc.wpart[warp][k] = w; }
            __syncthreads();
            if (tid < 6) { double a = 0.0; for (int w = 0; w < LG_WARPS; ++w) a += sc.wpart[w][tid]; sc.slot1[par][tid] = a; }
            cluster.sync();
            if (tid < 6) {
                double v[LG_CLUSTER];
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) v[r] = cluster.map_shared_rank(&sc.slot1[par][0], r)[tid];
                double a = 0.0;
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) a += v[r];
                sc.bcast[tid] = a * inv_n;
            }
            __syncthreads();
            double cen[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) cen[k] = sc.bcast[k];
            // ---- pass 2: mean distance to the centroid (:35) ----------------------------------------
            double d[3] = {0, 0, 0};
            for (int i = tid; i < npts; i += LG_THREADS) {
                const double2* q = reinterpret_cast<const double2*>(pts + 6 * i);
                const double2 a = q[0], b = q[1], c = q[2];
                double dx = a.x - cen[0], dy = a.y - cen[1]; d[0] += sqrt(dx * dx + dy * dy);
                dx = b.x - cen[2]; dy = b.y - cen[3]; d[1] += sqrt(dx * dx + dy * dy);
                dx = c.x - cen[4]; dy = c.y - cen[5]; d[2] += sqrt(dx * dx + dy * dy);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double w = warp_sum_l(d[k]); if (lane == 0) sc.wpart[warp][8 + k] = w; }
            __syncthreads();
            if (tid < 3) { double a = 0.0; for (int w = 0; w < LG_WARPS; ++w) a += sc.wpart[w][8 + tid]; sc.slot2[par][tid] = a; }
            cluster.sync();
            if (tid < 3) {
                double v[LG_CLUSTER];
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) v[r] = cluster.map_shared_rank(&sc.slot2[par][0], r)[tid];
                double a = 0.0;
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) a += v[r];
                sc.bcast[8 + tid] = 1.4142135623730951 / (a * inv_n);                          // :36
            }
            __syncthreads();
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                s[v] = sc.bcast[8 + v];
                t[2 * v] = -s[v] * cen[2 * v]; t[2 * v + 1] = -s[v] * cen[2 * v + 1];         // :37
            }
        }
        // ---- pass 3: 96 moments.  Warp w handles the 24 moments with view-3 feature index beta = w
        //      (all 6 view-1 features x 4 view-2 features) over all points of the slice. -----------------------
        const int beta = warp & 3;
        double acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.0;
        for (int i = lane; i < npts; i += 32) {
            const double2* q = reinterpret_cast<const double2*>(pts + 6 * i);
            const double2 a = q[0], b = q[1], c = q[2];
            const double x1 = s[0] * a.x + t[0], y1 = s[0] * a.y + t[1];
            const double x2 = s[1] * b.x + t[2], y2 = s[1] * b.y + t[3];
            const double x3 = s[2] * c.x + t[4], y3 = s[2] * c.y + t[5];
            const double m3 = (beta == 0) ? 1.0 : ((beta == 1) ? -x3 : ((beta == 2) ? -y3 : x3 * x3 + y3 * y3));
            const double tg[4] = {m3, -m3 * x2, -m3 * y2, m3 * (x2 * x2 + y2 * y2)};
            const double a6[5] = {x1 * x1, x1 * y1, x1, y1 * y1, y1};
#pragma unroll
            for (int al = 0; al < 5; ++al)
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[al * 4 + g] = fma(a6[al], tg[g], acc[al * 4 + g]);
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[20 + g] += tg[g];
        }
        const double tot = warp_reduce_transposed32(acc, lane);     // lane L (< 24) now holds moment (alpha = L/4, gamma = L%4)
        sc.wpart[warp][lane] = tot;
        __syncthreads();
        if (tid < 96) {
            const int al = tid >> 4, be = (tid >> 2) & 3, ga = tid & 3;         // moment index = alpha*16 + beta*4 + gamma
            sc.slot3[par][tid] = sc.wpart[be][al * 4 + ga];
        }
        cluster.sync();
        if (rank == 0) {
            double* rec = ws + scene * CORE_WS_TFT;
            if (tid < 96) {
                double v[LG_CLUSTER];
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) v[r] = cluster.map_shared_rank(&sc.slot3[par][0], r)[tid];
                double a = 0.0;
#pragma unroll
                for (int r = 0; r < LG_CLUSTER; ++r) a += v[r];
                rec[CW_MOM_L + tid] = a;
            }
            if (tid < 9) rec[CW_STATS_L + tid] = (tid < 3) ? s[tid] : t[tid - 3];
        }
        // No trailing cluster barrier: the reduction slots alternate with the scene parity, so a CTA that runs
        // ahead writes the other copy; it cannot lap a reader by two scenes because three cluster barriers of
        // the scene in between separate them.  The shared-memory slice is private to this CTA and every thread
        // passed the barrier above after its last read of it.
    }
    cluster.sync();      // keep every CTA's shared memory alive until all remote reads are done
}

// returns 0 when the shape is not supported by this kernel (caller falls back to tft_stage1_kernel)
int launch_tft_moments_large(const double* corresp, int n, long long B, int normalize, double* ws, int sm_count,
                             cudaStream_t stream) {
    if (B <= 0) return 1;
    const int slice_pts = (n + LG_CLUSTER - 1) / LG_CLUSTER;
    const size_t smem = (size_t)slice_pts * 48 + sizeof(LargeScratch) + 128;
    if (smem > 200 * 1024) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tft_moments_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return 0;
        attr_set = true;
    }
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    long long clusters = ((long long)sm_count * per_sm) / LG_CLUSTER;
    if (clusters > B) clusters = B;
    if (clusters < 1) clusters = 1;
    tft_moments_large_kernel<<<(unsigned)(clusters * LG_CLUSTER), LG_THREADS, smem, stream>>>(corresp, n, B, slice_pts,
                                                                                            normalize, ws);
    return 1;
}

}  // namespace tvf
