/* mex.h -- STUB of the MATLAB MEX C API, just large enough to compile and exercise the gateways in
 * mex/ without MATLAB (neither MATLAB nor Octave exists in the build image).  Under a real MATLAB
 * (`mex -I../include -L../tft_vs_fund_b200 -ltvf linearTFT.cpp`) or Octave (`mkoctfile --mex ...`)
 * the vendor's own mex.h is used instead and this file is ignored. */
#ifndef TVF_MEX_STUB_H_
#define TVF_MEX_STUB_H_
#include <stddef.h>
#ifdef __cplusplus
#include <stdexcept>
#include <string>
extern "C" {
#endif

typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;

typedef struct mxArray_tag {
    int ndim;
    mwSize dims[4];
    double* data;
    int is_double, is_complex, is_sparse;
} mxArray;

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
void mxDestroyArray(mxArray* a);
void* mxCalloc(size_t n, size_t size);       /* MATLAB frees these automatically when the MEX function leaves, also through an error */
void mxFree(void* p);
double* mxGetPr(const mxArray* a);
mwSize mxGetM(const mxArray* a);
mwSize mxGetN(const mxArray* a);            /* product of dims 2..end, as in MATLAB */
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsSparse(const mxArray* a);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);   /* stub: throws mex_stub_error */
/* the gateway entry point has C linkage, exactly as the vendor header declares it */
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
void mexLock(void);
int mexAtExit(void (*fn)(void));

#ifdef __cplusplus
}
struct mex_stub_error : public std::runtime_error {
    std::string id;
    mex_stub_error(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};
#endif
#endif
