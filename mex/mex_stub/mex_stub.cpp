// Implementation of the stub mx*/mex* API (test infrastructure for the gateways, not product code).
#include "mex.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {

mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID, mxComplexity) {
    mxArray* a = (mxArray*)std::calloc(1, sizeof(mxArray));
    a->ndim = (int)(ndim < 2 ? 2 : ndim);
    size_t n = 1;
    for (int i = 0; i < 4; ++i) { a->dims[i] = (i < (int)ndim) ? dims[i] : 1; n *= a->dims[i]; }
    a->data = (double*)std::calloc(n ? n : 1, sizeof(double));
    a->is_double = 1;
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    const mwSize d[2] = {m, n};
    return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, c);
}
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); a->data[0] = v; return a; }
void* mxCalloc(size_t n, size_t size) { return std::calloc(n ? n : 1, size ? size : 1); }
void mxFree(void* p) { std::free(p); }
void mxDestroyArray(mxArray* a) { if (a) { std::free(a->data); std::free(a); } }
double* mxGetPr(const mxArray* a) { return a->data; }
mwSize mxGetM(const mxArray* a) { return a->dims[0]; }
mwSize mxGetN(const mxArray* a) { mwSize n = 1; for (int i = 1; i < a->ndim; ++i) n *= a->dims[i]; return n; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return (mwSize)a->ndim; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims; }
int mxIsDouble(const mxArray* a) { return a->is_double; }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsSparse(const mxArray* a) { return a->is_sparse; }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw mex_stub_error(id ? id : "", buf);
}
void mexLock(void) {}
int mexAtExit(void (*)(void)) { return 0; }

}  // extern "C"
