// MEX gateway: [R_t_2,R_t_3,Reconst,T,iter] = LinearFPoseEstimation(Corresp,CalM)
// drop-in for F_methods/LinearFPoseEstimation.m:1.
#include "tvf_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    tvf_mex::pose_gateway(nlhs, plhs, nrhs, prhs, "LinearFPoseEstimation",
                          [](tvf_handle_t h, const double* c, const double* k, int kb, int n, int64_t B, double* Rt2,
                             double* Rt3, double* rec, double* T, int32_t* st, int32_t*) {
                              return tvf_linear_f_pose(h, c, k, kb, n, B, Rt2, Rt3, rec, T, nullptr, nullptr, nullptr, st);
                          });
}
