// Development probe: which thread-block-cluster shapes fit on this GPU (how many co-resident clusters per shared-memory
// footprint), and whether non-power-of-two cluster sizes launch.  nvcc -arch=sm_100a -o tools/_build/probe_clusters tools/probe_clusters.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void probe_kernel(int* out) {
    extern __shared__ unsigned char smem[];
    cg::cluster_group c = cg::this_cluster();
    if (threadIdx.x == 0) { smem[0] = 1; atomicAdd(out, (int)c.num_blocks()); }
    c.sync();
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs %d smem/SM %zu smem/block optin %zu\n", p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16};
    const int smems[] = {30, 60, 66, 72, 90, 100, 110, 130, 170, 210};
    const int threads[] = {128, 160, 256, 416};
    for (int t : threads)
        for (int sm : smems) {
            printf("threads %3d smem %3d KB: ", t, sm);
            for (int cs : sizes) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(t); cfg.dynamicSmemBytes = (size_t)sm * 1024;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int nc = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, probe_kernel, &cfg);
                if (e != cudaSuccess) { cudaGetLastError(); printf("c%d:err ", cs); continue; }
                printf("c%d:%d(%d) ", cs, nc, nc * cs);
            }
            printf("\n");
        }
    // launch test for a few sizes
    int* d; cudaMalloc(&d, 4);
    for (int cs : {3, 5, 6, 7, 12, 16}) {
        cudaMemset(d, 0, 4);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 4); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 60 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d);
        cudaError_t e2 = cudaDeviceSynchronize();
        int v = 0; cudaMemcpy(&v, d, 4, cudaMemcpyDeviceToHost);
        printf("launch cluster %d: %s / %s -> sum of num_blocks %d (expect %d)\n", cs, cudaGetErrorString(e), cudaGetErrorString(e2), v, cs * cs * 4);
        cudaGetLastError();
    }
    return 0;
}
