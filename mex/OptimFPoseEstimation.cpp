// MEX gateway: [R_t_2,R_t_3,Reconst,T,iter] = OptimFPoseEstimation(Corresp,CalM)
// drop-in for F_methods/OptimFPoseEstimation.m:1 (method 8 of example.m:32-40 / experiments.m:51-59);
// iter = it1 + it2, the Gauss-Helmert iterations of the two optimF calls (:47-49).
#include "tvf_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    tvf_mex::pose_gateway(nlhs, plhs, nrhs, prhs, "OptimFPoseEstimation",
                          [](tvf_handle_t h, const double* c, const double* k, int kb, int n, int64_t B, double* Rt2,
                             double* Rt3, double* rec, double* T, int32_t* st, int32_t* it) {
                              return tvf_optim_f_pose(h, c, k, kb, n, B, Rt2, Rt3, rec, T, nullptr, nullptr, nullptr, it, st);
                          });
}
