// MEX gateway: [R_t_2,R_t_3,Reconst,T,iter] = LinearTFTPoseEstimation(Corresp,CalM)
// drop-in for TFT_methods/LinearTFTPoseEstimation.m:1 (a MEX file of the same base name earlier on the
// MATLAB path shadows the .m), so experiments.m:108 / example.m:42 call it unchanged.
#include "tvf_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    tvf_mex::pose_gateway(nlhs, plhs, nrhs, prhs, "LinearTFTPoseEstimation",
                          [](tvf_handle_t h, const double* c, const double* k, int kb, int n, int64_t B, double* Rt2,
                             double* Rt3, double* rec, double* T, int32_t* st, int32_t*) {
                              return tvf_linear_tft_pose(h, c, k, kb, n, B, Rt2, Rt3, rec, T, nullptr, st);
                          });
}
