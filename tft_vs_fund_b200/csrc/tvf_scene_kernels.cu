// tvf_scene_kernels.cu -- batched trials of the synthetic sweep generated on the device.
// One thread per trial; the MT19937 state and the 6 x (n+100) coordinate scratch live in local memory.
#include "tvf_kernels.h"
#include "tvf_scene.cuh"

namespace tvf {

__global__ void __launch_bounds__(64)
sweep_trials_kernel(long long first_trial, long long B, int n, const double* __restrict__ noise_levels, int L,
                    const double* __restrict__ P, double hi_x, double hi_y, double* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long j = first_trial + b;                    // global trial index (experiments.m:74-95)
    const uint32_t seed = (uint32_t)(j / L + 1);
    const double noise = noise_levels[j % L];
    MT19937 rng;
    double c[6 * SCENE_MAX_POINTS];
    unsigned char arr[SCENE_MAX_POINTS];
    signed char outpos[SCENE_MAX_POINTS];
    double Pl[36];
    for (int i = 0; i < 36; ++i) Pl[i] = P[i];
    scene_trial(rng, Pl, n, noise, seed, hi_x, hi_y, out + b * 6 * n, c, arr, outpos);
}

void launch_sweep_trials(long long first_trial, long long B, int n, const double* d_noise_levels, int L, const double* d_P,
                         double hi_x, double hi_y, double* d_out, cudaStream_t stream) {
    if (B <= 0) return;
    sweep_trials_kernel<<<(unsigned)((B + 63) / 64), 64, 0, stream>>>(first_trial, B, n, d_noise_levels, L, d_P, hi_x, hi_y, d_out);
}

}  // namespace tvf
