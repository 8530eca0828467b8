// Development microbenchmark: dependent-issue latencies (cycles) of the instructions on the Gauss-Jordan sweep's serial
// chain, one warp per SM sub-partition, and the single-warp issue rate of independent DFMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_lat tools/microbench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
template <int MODE>
__global__ void lat_kernel(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    sm[lane] = seed + lane; sm[lane + 32] = seed;
    __syncwarp();
    double a = seed + lane * 1e-3, b = 1.0000001, c = 1e-9;
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    int idx = lane;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) {
        if (MODE == 0) a = fma(a, b, c);                                   // dependent DFMA
        if (MODE == 1) a = a * b;                                          // dependent DMUL
        if (MODE == 2) a = a + c;                                          // dependent DADD
        if (MODE == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r; }   // MUFU.RCP64H chain
        if (MODE == 4) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);   // dependent 64-bit shuffle (2 SHFL)
        if (MODE == 5) { idx = (int)sm[idx & 31] & 31; }                   // dependent LDS.64 (+ conversion)
        if (MODE == 6) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
                         x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c); }   // 8 independent DFMAs
        if (MODE == 7) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] + c; __syncwarp(); }           // STS -> sync -> LDS -> DADD round trip
        if (MODE == 8) { x0 = fma(x0, x1, x2); x3 = fma(x3, x4, x5); x6 = fma(x6, x7, x0); x1 = fma(x1, x2, x3);
                         x4 = fma(x4, x5, x6); x7 = fma(x7, x0, x1); x2 = fma(x2, x3, x4); x5 = fma(x5, x6, x7); }   // 3 fresh operands each
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + lane] = a + x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + idx;
}

// shared-memory delivery rate: `warps` warps per SM each issue independent LDS.128 / LDS.64 with a given address pattern.
// PATTERN 0: all lanes the same address (broadcast), 1: two addresses (one per half-warp), 2: every lane its own 16 bytes
template <int PATTERN, int WIDTH>
__global__ void lds_rate_kernel(double* out, long long* cyc) {
    __shared__ __align__(16) double sm[4096];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5;
    __syncthreads();
    const int sel = (PATTERN == 0) ? 0 : ((PATTERN == 1) ? (lane >> 4) * 32 : lane * 2);
    const double* base = sm + warp * 64 + sel;
    double acc = 0.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (WIDTH == 16) { const double2 v = *reinterpret_cast<const double2*>(base + ((k * 2 + it) & 30)); acc += v.x + v.y; }
            else { acc += base[(k + it) & 31]; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096 * 8);
    const char* names[] = {"dependent DFMA", "dependent DMUL", "dependent DADD", "dependent MUFU.RCP64H", "dependent SHFL.64",
                           "dependent LDS.64+F2I", "8 independent DFMA (per group)", "STS-sync-LDS-DADD round trip", "8 DFMA, 3 fresh operands"};
    for (int warps : {1, 2, 4}) {
        printf("-- %d warp(s) per CTA, 1 CTA (cycles per repetition) --\n", warps);
        for (int m = 0; m < 9; ++m) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (m) {
                    case 0: lat_kernel<0><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 1: lat_kernel<1><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 2: lat_kernel<2><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 3: lat_kernel<3><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 4: lat_kernel<4><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 5: lat_kernel<5><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 6: lat_kernel<6><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 7: lat_kernel<7><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                    case 8: lat_kernel<8><<<1, 32 * warps>>>(out, cyc, 1.5); break;
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("%-34s %.2f\n", names[m], (double)h / REP);
        }
    }
    printf("-- shared-memory delivery (SM cycles per LDS instruction, 16 warps on one SM, includes 2-3 DADD per load) --\n");
    {
        long long h;
        lds_rate_kernel<0, 16><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 broadcast      %.2f\n", (double)h / (64 * 16 * 16));
        lds_rate_kernel<1, 16><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 two addresses  %.2f\n", (double)h / (64 * 16 * 16));
        lds_rate_kernel<2, 16><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 per-lane       %.2f\n", (double)h / (64 * 16 * 16));
        lds_rate_kernel<0, 8><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.64 broadcast       %.2f\n", (double)h / (64 * 16 * 16));
        lds_rate_kernel<1, 8><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.64 two addresses   %.2f\n", (double)h / (64 * 16 * 16));
        lds_rate_kernel<2, 8><<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.64 per-lane        %.2f\n", (double)h / (64 * 16 * 16));
    }
    return 0;
}
