// tvf_kernels.h -- internal interface between the C-ABI layer (tvf_api.cu) and
// the kernel translation units.  Not installed; the public contract is include/tvf.h.
#pragma once
#include <cuda_runtime.h>

namespace tvf {

// Where a warp/thread finds the correspondences of problem b, point i.
//   packed != 0 : p1 is `Corresp` 6 x n x B (column-major; 48 contiguous bytes per point)
//   packed == 0 : p1,p2,p3 are rows x n x B each (rows = 2, or 3 for homogeneous input)
struct CoreInput {
    const double* p1;
    const double* p2;
    const double* p3;
    int packed;
    int rows;
    int n;
    long long B;
    int normalize;   // 1: *PoseEstimation path (Normalize2Ddata + undo); 0: estimator called directly
};

// per-problem work-space records (doubles) handed between the estimator stages
constexpr int CORE_WS_TFT = 140;   // 96 moments | 9 normalisation stats (+1) | t1 27 (+1) | e21,e31
constexpr int CORE_WS_F = 36;      // raw f (2 x 9) | outer stats 9 | inner stats 9

// returns 1 when only the moments half ran (split stage 1): the caller then launches launch_tft_stage1_solve
int launch_tft_stage1(const CoreInput& in, double* ws, int* status, int sm_count, cudaStream_t stream);
// large-n stage 1: cluster/TMA moments kernel (returns 0 if the shape is unsupported -> use launch_tft_stage1)
// followed by the solve-only half of stage 1
constexpr int LARGE_N_MIN = 1024;
int launch_tft_moments_large(const double* corresp, int n, long long B, int normalize, double* ws, int sm_count,
                             cudaStream_t stream);
void launch_tft_stage1_solve(long long B, double* ws, int* status, int sm_count, cudaStream_t stream);
void launch_tft_epipoles(double* ws, long long B, cudaStream_t stream);
void launch_tft_stage2(const CoreInput& in, const double* ws, double* T, double* P2, double* P3, int* status,
                       int sm_count, cudaStream_t stream);
void launch_f_stage1(const CoreInput& in, double* ws, int* status, int sm_count, cudaStream_t stream);
// normalize != 0: two pairs per problem (F21, F31 -> F[18*b]); else one (F[9*b])
void launch_f_finish(const double* ws, int normalize, long long B, double* F, cudaStream_t stream);

// ---- Gauss-Helmert refinement of F (optimF.m; tvf_gh_kernels.cu).  ws: f_stage1 records of the pose path
// (two pairs per problem); Fio: 18 x B, receives [F21 F31]; iters: 2 x B.  Returns 0 if n is too large.
int optimf_max_n();
int launch_optimf_gh(const double* corresp, int n, long long B, const double* ws, double* Fio, int* iters, int* status,
                     int sm_count, cudaStream_t stream);
void launch_sum_pairs(const int* in2, long long B, int* out, cudaStream_t stream);

// ---- pose tail -----------------------------------------------------------------------------
struct PoseTailArgs {
    const double* corresp;     // 6 x n x B
    const double* calm;        // 9 x 3 (shared) or 9 x 3 x B
    int calm_batched;
    int n;
    long long B;
    // workspace (device)
    double* cand;              // CAND_SIZE x B
    int* votes;                // 10 x B : 8 votes, nanmask pair 2, nanmask pair 3
    double* scale;             // 2 x B : num, den
    // outputs (device; any may be null)
    double* Rt2;               // 12 x B
    double* Rt3;               // 12 x B
    double* reconst;           // 3 x n x B
    double* repr_err;          // B
    int* status;               // B (or-ed into)
};

// mode 0: model = T (27 x B, pixel coordinates); mode 1: model = [F21 F31] (18 x B)
void launch_candidates(int mode, const double* model, const PoseTailArgs& a, cudaStream_t stream);
// TAIL_FUSED_MIN_N <= n <= TAIL_FUSED_MAX_N: votes + scale + final in one launch; otherwise the three kernels below
constexpr int TAIL_FUSED_MIN_N = 7;
constexpr int TAIL_FUSED_MAX_N = 256;
void launch_pose_tail_fused(const PoseTailArgs& a, int sm_count, cudaStream_t stream);
void launch_votes(const PoseTailArgs& a, int sm_count, cudaStream_t stream);
void launch_scale(const PoseTailArgs& a, int sm_count, cudaStream_t stream);
void launch_final(const PoseTailArgs& a, int sm_count, cudaStream_t stream);
// LinearFPoseEstimation.m:78  T = TFT_from_P(K1*eye(3,4), K2*R_t_2, K3*R_t_3)
void launch_tft_from_pose(const double* calm, int calm_batched, const double* Rt2, const double* Rt3,
                          long long B, double* T, cudaStream_t stream);

// ---- stand-alone reference functions (batched) ---------------------------------------------
void launch_normalize2d(const double* pts, int n, long long B, double* out, double* Nmat, cudaStream_t s);
void launch_transform_tft(const double* T, const double* M1, const double* M2, const double* M3, int mats_batched,
                          int inverse, long long B, double* Tout, cudaStream_t s);
void launch_tft_from_p(const double* P1, const double* P2, const double* P3, long long B, double* T, cudaStream_t s);
void launch_triangulate(const double* P, int M, int cams_batched, const double* pts, int rows, int n, long long B,
                        double* X, cudaStream_t s);
void launch_repr_error(const double* P, int M, int cams_batched, const double* corresp, int rows, int n, long long B,
                       const double* pts3d, int pts_rows, double* err, cudaStream_t s);
void launch_project3d(const double* pts3d, const double* P, int M, int cams_batched, int n, long long B, double* out,
                      cudaStream_t s);
void launch_ang_error(const double* Rt_true, int true_batched, const double* Rt_est, long long B, double* rot,
                      double* tr, cudaStream_t s);

// ---- per-noise-level evaluation sums (f3): partial is (L*Q) x 5, table L x 5 = [sum repr, sum rot, sum t, count, skipped]
void launch_sweep_eval_accumulate(const double* Rt2, const double* Rt3, const double* repr, const int* status,
                                  long long first_trial, long long B, int L, int Q, const double* d_Rt0, double* d_partial,
                                  cudaStream_t s);
void launch_sweep_eval_finish(const double* d_partial, int L, int Q, double* d_table, cudaStream_t s);

// ---- synthetic sweep trials generated on the device (f1) -----------------------------------------------
constexpr int SWEEP_MAX_N = 60;      // n + 100 points per scene must fit the per-thread scratch
void launch_sweep_trials(long long first_trial, long long B, int n, const double* d_noise_levels, int L, const double* d_P,
                         double hi_x, double hi_y, double* d_out, cudaStream_t stream);

}  // namespace tvf
