/* Minimal stand-in for MATLAB's mex.h -- TEST INFRASTRUCTURE.  Just enough of the mx/mex API
 * (as documented by MathWorks) for tests/test_mex_gateway.py to compile
 * optimal-control-solvers_b200/matlab/ocs_b200_mex.c and drive mexFunction from a C program. */
#ifndef OCS_STUB_MEX_H
#define OCS_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>

typedef size_t mwSize;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15, mxCHAR_CLASS = 4 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

typedef struct mxArray_tag {
    mxClassID cls;
    mwSize ndim;
    mwSize dims[4];
    void *data;
} mxArray;

#ifdef __cplusplus
extern "C" {
#endif
int mxIsUint64(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxIsComplex(const mxArray *a);
int mxIsEmpty(const mxArray *a);
mxClassID mxGetClassID(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
void *mxGetData(const mxArray *a);
double *mxGetPr(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
char *mxArrayToString(const mxArray *a);
void mxFree(void *p);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateString(const char *s);
void mxDestroyArray(mxArray *a);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexLock(void);
void mexUnlock(void);
int mexIsLocked(void);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
