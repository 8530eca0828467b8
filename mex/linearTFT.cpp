// MEX gateway: [T,P1,P2,P3] = linearTFT(p1,p2,p3)   drop-in for TFT_methods/linearTFT.m:1
// (called with one output by LinearTFTPoseEstimation.m:50 and with four by the Gauss-Helmert methods,
// e.g. ResslTFTPoseEstimation.m:53).  p* are 2xN or 3xN (or ...xB).
#include "tvf_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    using namespace tvf_mex;
    if (nrhs != 3) mexErrMsgIdAndTxt("TFT_vs_Fund:nargin", "linearTFT(p1,p2,p3) takes three inputs");
    if (nlhs > 4) mexErrMsgIdAndTxt("TFT_vs_Fund:nargout", "linearTFT returns at most four outputs");
    for (int i = 0; i < 3; ++i) require_real_double(prhs[i], "p");
    const Dims d = dims3(prhs[0]);
    for (int i = 1; i < 3; ++i) {
        const Dims e = dims3(prhs[i]);
        if (e.rows != d.rows || e.n != d.n || e.B != d.B) mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "p1,p2,p3 must have equal size");
    }
    if (d.rows != 2 && d.rows != 3) mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "points must be 2xN or 3xN");
    tvf_handle_t h = handle();
    mxArray* T = make_tensor(d.B);
    mxArray* P[3] = {make(3, 4, d.B), make(3, 4, d.B), make(3, 4, d.B)};
    const int rc = tvf_linear_tft(h, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), (int)d.rows, (int)d.n,
                                  (int64_t)d.B, mxGetPr(T), mxGetPr(P[1]), mxGetPr(P[2]), nullptr);
    if (rc < 0) { mxDestroyArray(T); for (auto* p : P) mxDestroyArray(p); check(rc, h); }
    for (mwSize b = 0; b < d.B; ++b) { double* p1 = mxGetPr(P[0]) + 12 * b; p1[0] = p1[4] = p1[8] = 1.0; }   // P1=eye(3,4) (:88)
    mxArray* outs[4] = {T, P[0], P[1], P[2]};
    const int nout = nlhs < 1 ? 1 : nlhs;
    for (int i = 0; i < 4; ++i) { if (i < nout) plhs[i] = outs[i]; else mxDestroyArray(outs[i]); }
}
