// MEX gateway: [F,iter] = optimF(p1,p2)   drop-in for F_methods/optimF.m:1.
#include "tvf_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    using namespace tvf_mex;
    if (nrhs != 2) mexErrMsgIdAndTxt("TFT_vs_Fund:nargin", "optimF(p1,p2) takes two inputs");
    if (nlhs > 2) mexErrMsgIdAndTxt("TFT_vs_Fund:nargout", "optimF returns at most two outputs");
    require_real_double(prhs[0], "p1"); require_real_double(prhs[1], "p2");
    const Dims a = dims3(prhs[0]), b = dims3(prhs[1]);
    if (a.n != b.n || a.n < 8) mexErrMsgIdAndTxt("TFT_vs_Fund:linearF", TVF_LINEARF_ERRMSG);        // optimF.m:36-38
    if (a.rows != b.rows || a.B != b.B || (a.rows != 2 && a.rows != 3))
        mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "p1,p2 must both be 2xN or 3xN");
    tvf_handle_t h = handle();
    mxArray* F = make(3, 3, a.B);
    int32_t* it = temp_i32(a.B);              // mxCalloc: released by MATLAB on the error paths below
    const int rc = tvf_optim_f(h, mxGetPr(prhs[0]), mxGetPr(prhs[1]), (int)a.rows, (int)a.n, (int64_t)a.B, mxGetPr(F), it, nullptr);
    if (rc < 0) { mxDestroyArray(F); check(rc, h); }
    plhs[0] = F;
    if (nlhs > 1) {
        plhs[1] = (a.B == 1) ? mxCreateDoubleScalar((double)it[0]) : make(a.B, 1, 1);
        if (a.B != 1) for (mwSize k = 0; k < a.B; ++k) mxGetPr(plhs[1])[k] = (double)it[k];
    }
    mxFree(it);
}
