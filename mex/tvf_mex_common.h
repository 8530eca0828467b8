// tvf_mex_common.h -- shared plumbing of the MEX gateways: one persistent libtvf handle per MATLAB
// process (mexLock + mexAtExit), argument checks, batched shapes.
//
// Contract: inputs are borrowed `const mxArray*` (real, full, double); outputs are created
// with mxCreate* and handed to MATLAB; plhs[0] is always set; errors leave through mexErrMsgIdAndTxt after
// nothing of ours is left allocated (all device memory lives in the persistent handle).  MATLAB calls MEX
// on its interpreter thread only, so every mx* call happens on the calling thread; CUDA work runs on the
// handle's private streams and is synchronised before the gateway returns.
//
// Devices: the environment variable TVF_DEVICES ("0", "0,1,2,3", "all"; default "0") selects the GPUs of the
// persistent handle; with more than one the batched pose calls are sharded over them inside libtvf
// (tvf_create_multi: one host thread per device, experiments.m:91-143 is the loop that fans out).
// TVF_HOST_REGISTER=1 page-locks large mxArray buffers for the duration of a call (tvf_set_host_register).
// Temporaries come from mxCalloc: MATLAB releases them itself when mexErrMsgIdAndTxt leaves the function.
//
// Batched superset: a trailing dimension B on the inputs (Corresp 6xNxB, p 2xNxB, CalM 9x3 or 9x3xB)
// returns 3x4xB, 3xNxB, 3x3x3xB, ... ; B = 1 is exactly the reference signature.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "mex.h"
#include "tvf.h"

namespace tvf_mex {

inline tvf_handle_t& handle_ref() { static tvf_handle_t h = nullptr; return h; }

inline void cleanup() {
    if (handle_ref()) { tvf_destroy(handle_ref()); handle_ref() = nullptr; }
}

inline tvf_handle_t handle() {
    tvf_handle_t& h = handle_ref();
    if (!h) {
        int devs[64]; int nd = 0;
        const char* env = std::getenv("TVF_DEVICES");
        if (env && std::strcmp(env, "all") == 0) {
            const int cnt = tvf_device_count();
            for (int i = 0; i < cnt && i < 64; ++i) devs[nd++] = i;
        } else if (env && *env) {
            for (const char* c = env; *c && nd < 64;) {
                char* end = nullptr;
                const long v = std::strtol(c, &end, 10);
                if (end == c) break;
                devs[nd++] = (int)v;
                c = end;
                while (*c == ',' || *c == ' ') ++c;
            }
        }
        if (nd == 0) devs[nd++] = 0;
        if (tvf_create_multi(&h, devs, nd) != TVF_OK)
            mexErrMsgIdAndTxt("TFT_vs_Fund:noDevice", "libtvf: %s", tvf_last_error(nullptr));
        const char* reg = std::getenv("TVF_HOST_REGISTER");
        if (reg && *reg && *reg != '0') tvf_set_host_register(h, 1);
        mexLock();
        mexAtExit(cleanup);
    }
    return h;
}

inline int32_t* temp_i32(mwSize count) { return static_cast<int32_t*>(mxCalloc(count ? count : 1, sizeof(int32_t))); }

inline void require_real_double(const mxArray* a, const char* name) {
    if (!mxIsDouble(a) || mxIsComplex(a) || mxIsSparse(a))
        mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "%s must be a real, full, double array", name);
}

// rows x n x B view of an input (B = product of the trailing dimensions beyond the second)
struct Dims { mwSize rows, n, B; };
inline Dims dims3(const mxArray* a) {
    const mwSize nd = mxGetNumberOfDimensions(a);
    const mwSize* d = mxGetDimensions(a);
    Dims r{d[0], nd > 1 ? d[1] : 1, 1};
    for (mwSize i = 2; i < nd; ++i) r.B *= d[i];
    return r;
}

inline mxArray* make(mwSize a, mwSize b, mwSize B) {
    if (B == 1) return mxCreateDoubleMatrix(a, b, mxREAL);
    const mwSize d[3] = {a, b, B};
    return mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
}

inline mxArray* make_tensor(mwSize B) {
    const mwSize d[4] = {3, 3, 3, B};
    return mxCreateNumericArray(B == 1 ? 3 : 4, d, mxDOUBLE_CLASS, mxREAL);
}

inline void check(int rc, tvf_handle_t h) {
    if (rc == TVF_ERR_TOO_FEW_POINTS)     // linearF.m:35-37, same text
        mexErrMsgIdAndTxt("TFT_vs_Fund:linearF", TVF_LINEARF_ERRMSG);
    if (rc < 0) mexErrMsgIdAndTxt("TFT_vs_Fund:runtime", "libtvf error %d: %s", rc, tvf_last_error(h));
}

// [R_t_2,R_t_3,Reconst,T,iter] = Method(Corresp,CalM) for the pose-estimation methods
template <typename Call>
inline void pose_gateway(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[], const char* name, Call call) {
    if (nrhs != 2) mexErrMsgIdAndTxt("TFT_vs_Fund:nargin", "%s(Corresp,CalM) takes two inputs", name);
    if (nlhs > 5) mexErrMsgIdAndTxt("TFT_vs_Fund:nargout", "%s returns at most five outputs", name);
    require_real_double(prhs[0], "Corresp"); require_real_double(prhs[1], "CalM");
    const Dims c = dims3(prhs[0]), k = dims3(prhs[1]);
    if (c.rows != 6) mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "Corresp must be 6xN (or 6xNxB)");
    if (k.rows != 9 || k.n != 3 || (k.B != 1 && k.B != c.B))
        mexErrMsgIdAndTxt("TFT_vs_Fund:badInput", "CalM must be 9x3 (or 9x3xB)");
    tvf_handle_t h = handle();
    mxArray* Rt2 = make(3, 4, c.B); mxArray* Rt3 = make(3, 4, c.B);
    mxArray* Rec = make(3, c.n, c.B); mxArray* T = make_tensor(c.B);
    int32_t* status = temp_i32(c.B);          // mxCalloc: released by MATLAB even when an error leaves this function
    int32_t* iters = temp_i32(c.B);
    const int rc = call(h, mxGetPr(prhs[0]), mxGetPr(prhs[1]), k.B != 1, (int)c.n, (int64_t)c.B, mxGetPr(Rt2),
                        mxGetPr(Rt3), mxGetPr(Rec), mxGetPr(T), status, iters);
    mxArray* outs[5] = {Rt2, Rt3, Rec, T, nullptr};
    if (rc < 0 || (c.B == 1 && (status[0] & (TVF_ST_NO_POSE_2 | TVF_ST_NO_POSE_3)))) {
        for (int i = 0; i < 4; ++i) mxDestroyArray(outs[i]);
        check(rc, h);
        // the reference stops here with "Undefined function or variable 'R_f'" (R_t_from_TFT.m:91-104)
        mexErrMsgIdAndTxt("TFT_vs_Fund:undefinedPose", "recover_R_t: no candidate pose received a non-negative vote");
    }
    // iter: the constant 0 of the linear methods (:62 / :77), it1+it2 for OptimFPoseEstimation (:49)
    outs[4] = (c.B == 1) ? mxCreateDoubleScalar((double)iters[0]) : make(c.B, 1, 1);
    if (c.B != 1) for (mwSize b = 0; b < c.B; ++b) mxGetPr(outs[4])[b] = (double)iters[b];
    const int nout = nlhs < 1 ? 1 : nlhs;
    for (int i = 0; i < 5; ++i) {
        if (i < nout) plhs[i] = outs[i]; else mxDestroyArray(outs[i]);
    }
    mxFree(status); mxFree(iters);
}

}  // namespace tvf_mex
