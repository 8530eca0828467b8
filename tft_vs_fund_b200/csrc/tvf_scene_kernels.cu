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

// One thread per SEED: the L noise levels of a seed share the seeding, the sub-sample permutation and the whole
// first pass of the generator (scene_seed_levels), ~L times less work than one thread per trial.
__global__ void __launch_bounds__(64)
sweep_seeds_kernel(long long first_trial, long long B, int n, const double* __restrict__ noise_levels, int L,
                   const double* __restrict__ P, double hi_x, double hi_y, double* __restrict__ out) {
    const long long s0 = first_trial / L;
    const long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;       // 0-based seed index, seed = s + 1
    const long long j_lo = s * L, last = first_trial + B;
    if (j_lo >= last) return;
    const int lv_lo = (int)((first_trial > j_lo) ? first_trial - j_lo : 0);
    const int lv_hi = (int)((last - j_lo < L) ? last - j_lo : L);
    MT19937 rng, snap;
    double clean[6 * SCENE_MAX_POINTS], z[6 * SCENE_MAX_POINTS], c[6 * SCENE_MAX_POINTS];
    unsigned char arr[SCENE_MAX_POINTS];
    signed char outpos[SCENE_MAX_POINTS];
    double Pl[36];
    for (int i = 0; i < 36; ++i) Pl[i] = P[i];
    scene_seed_levels(rng, snap, Pl, n, noise_levels, lv_lo, lv_hi, (uint32_t)(s + 1), hi_x, hi_y,
                      out + (j_lo + lv_lo - first_trial) * 6 * n, clean, z, c, arr, outpos);
}

void launch_sweep_trials(long long first_trial, long long B, int n, const double* d_noise_levels, int L, const double* d_P,
                         double hi_x, double hi_y, double* d_out, cudaStream_t stream) {
    if (B <= 0) return;
    if (L >= 4) {
        const long long seeds = (first_trial + B - 1) / L - first_trial / L + 1;
        sweep_seeds_kernel<<<(unsigned)((seeds + 63) / 64), 64, 0, stream>>>(first_trial, B, n, d_noise_levels, L, d_P, hi_x, hi_y, d_out);
        return;
    }
    sweep_trials_kernel<<<(unsigned)((B + 63) / 64), 64, 0, stream>>>(first_trial, B, n, d_noise_levels, L, d_P, hi_x, hi_y, d_out);
}

}  // namespace tvf
