// microbench_fp64.cu -- which FP64 resources does a B200 SM have, and do they add up?
//   (1) DFMA only, (2) DMMA m8n8k4 only, (3) DMMA m16n8k8 only, (4) DFMA + DMMA interleaved in one warp,
//   (5) DFMA warps and DMMA warps side by side, (6) FP64 sqrt / rsqrt-seeded sqrt / division rates.
// Decides the large-n Gram design: tensor-core Gram only pays if DMMA is a pipe of its own.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double seed) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x + k;
    const double m = 0.999999, c = 1e-6;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    double r = 0; for (int k = 0; k < 8; ++k) r += a[k];
    if (r == 123.456) out[0] = r;
}
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double seed) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) { c[k][0] = seed + k; c[k][1] = seed - k; }
    const double a = 1e-3 * threadIdx.x, b = 1e-3;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
    double r = 0; for (int k = 0; k < 8; ++k) r += c[k][0] + c[k][1];
    if (r == 123.456) out[0] = r;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double seed) {
    double c[4][4];
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) c[k][j] = seed + k + j;
    const double a[4] = {1e-3 * threadIdx.x, 2e-3, 3e-3, 4e-3}, b[2] = {1e-3, 2e-3};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma1688(c[k], a, b);
    double r = 0; for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) r += c[k][j];
    if (r == 123.456) out[0] = r;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double seed) {
    double c[4][4];
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) c[k][j] = seed + k + j;
    const double a[8] = {1e-3 * threadIdx.x, 2e-3, 3e-3, 4e-3, 5e-3, 6e-3, 7e-3, 8e-3}, b[4] = {1e-3, 2e-3, 3e-3, 4e-3};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma16816(c[k], a, b);
    double r = 0; for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) r += c[k][j];
    if (r == 123.456) out[0] = r;
}
// one warp issues both: per iteration 4 DMMA m8n8k4 (4*256 FMA) + NF DFMA instructions (NF*32 FMA)
template <int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double seed) {
    double c[4][2], f[NF];
    for (int k = 0; k < 4; ++k) { c[k][0] = seed + k; c[k][1] = seed - k; }
    for (int k = 0; k < NF; ++k) f[k] = seed + threadIdx.x + k;
    const double a = 1e-3 * threadIdx.x, b = 1e-3, m = 0.999999, cc = 1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma884(c[k][0], c[k][1], a, b);
#pragma unroll
        for (int k = 0; k < NF; ++k) f[k] = fma(f[k], m, cc);
    }
    double r = 0; for (int k = 0; k < 4; ++k) r += c[k][0] + c[k][1];
    for (int k = 0; k < NF; ++k) r += f[k];
    if (r == 123.456) out[0] = r;
}
// half the warps of a CTA do DFMA, the other half DMMA
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double seed) {
    const int warp = threadIdx.x >> 5;
    double r = 0;
    if (warp & 1) {
        double a[8];
        for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x + k;
        const double m = 0.999999, c = 1e-6;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
        for (int k = 0; k < 8; ++k) r += a[k];
    } else {
        double c[8][2];
        for (int k = 0; k < 8; ++k) { c[k][0] = seed + k; c[k][1] = seed - k; }
        const double a = 1e-3 * threadIdx.x, b = 1e-3;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
        for (int k = 0; k < 8; ++k) r += c[k][0] + c[k][1];
    }
    if (r == 123.456) out[0] = r;
}
// special functions: results/s
__global__ void __launch_bounds__(256) k_sqrt(double* out, int iters, double seed) {
    double a[4];
    for (int k = 0; k < 4; ++k) a[k] = seed + threadIdx.x + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = sqrt(a[k] + 3.0);
    double r = 0; for (int k = 0; k < 4; ++k) r += a[k];
    if (r == 123.456) out[0] = r;
}
__global__ void __launch_bounds__(256) k_div(double* out, int iters, double seed) {
    double a[4];
    for (int k = 0; k < 4; ++k) a[k] = seed + threadIdx.x + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = 3.0 / (a[k] + 1.5);
    double r = 0; for (int k = 0; k < 4; ++k) r += a[k];
    if (r == 123.456) out[0] = r;
}
__device__ __forceinline__ double fast_sqrt(double x) {   // MUFU.RSQ64H seed + 2 Newton steps + residual correction
    int hi = __double2hiint(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    (void)hi;
    double h = 0.5 * y, g = x * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}
__global__ void __launch_bounds__(256) k_fsqrt(double* out, int iters, double seed) {
    double a[4];
    for (int k = 0; k < 4; ++k) a[k] = seed + threadIdx.x + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = fast_sqrt(a[k] + 3.0);
    double r = 0; for (int k = 0; k < 4; ++k) r += a[k];
    if (r == 123.456) out[0] = r;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, T = 256, iters = 1 << 14;
    double* out; cudaMalloc(&out, 64);
    const double thr = (double)blocks * T;
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    double ms;
    ms = time_ms([&] { k_dfma<<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"dfma_tflops\": %.2f", thr * iters * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma884<<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"dmma_m8n8k4_tflops\": %.2f", thr / 32 * iters * 8 * 512 / ms / 1e9);
    ms = time_ms([&] { k_dmma1688<<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"dmma_m16n8k8_tflops\": %.2f", thr / 32 * iters * 4 * 2048 / ms / 1e9);
    ms = time_ms([&] { k_dmma16816<<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"dmma_m16n8k16_tflops\": %.2f", thr / 32 * iters * 4 * 4096 / ms / 1e9);
    ms = time_ms([&] { k_mixed<8><<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"mixed_4dmma_8dfma\": {\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", thr / 32 * iters * 4 * 512 / ms / 1e9,
           thr * iters * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_mixed<16><<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"mixed_4dmma_16dfma\": {\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", thr / 32 * iters * 4 * 512 / ms / 1e9,
           thr * iters * 16 * 2 / ms / 1e9);
    ms = time_ms([&] { k_mixed<32><<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"mixed_4dmma_32dfma\": {\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", thr / 32 * iters * 4 * 512 / ms / 1e9,
           thr * iters * 32 * 2 / ms / 1e9);
    ms = time_ms([&] { k_split<<<blocks, T>>>(out, iters, 1.0); });
    printf(", \"split_warps\": {\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", thr / 64 * iters * 8 * 512 / ms / 1e9,
           thr / 2 * iters * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_sqrt<<<blocks, T>>>(out, iters / 8, 1.0); });
    printf(", \"dsqrt_gops\": %.2f", thr * (iters / 8) * 4 / ms / 1e6);
    ms = time_ms([&] { k_fsqrt<<<blocks, T>>>(out, iters / 8, 1.0); });
    printf(", \"fast_sqrt_gops\": %.2f", thr * (iters / 8) * 4 / ms / 1e6);
    ms = time_ms([&] { k_div<<<blocks, T>>>(out, iters / 8, 1.0); });
    printf(", \"ddiv_gops\": %.2f", thr * (iters / 8) * 4 / ms / 1e6);
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
