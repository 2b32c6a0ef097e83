/* Minimal stand-in for MATLAB's mex.h -- ONLY so that pcreg_b200/csrc/pcreg_mex.cpp can be type-checked
 * and exercised in CI without MATLAB (probed: no matlab/octave/mkoctfile/mex.h in this image).  The real
 * gateway is compiled with `mex` against MathWorks' header; nothing here ships. */
#ifndef FAKE_MEX_H
#define FAKE_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7,
               mxINT32_CLASS = 12, mxUINT32_CLASS = 13, mxINT64_CLASS = 14, mxUINT64_CLASS = 15, mxSTRUCT_CLASS = 2 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
void* mxGetData(const mxArray* a);
double* mxGetPr(const mxArray* a);
double mxGetScalar(const mxArray* a);
mwSize mxGetM(const mxArray* a);
mwSize mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
mxClassID mxGetClassID(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsSingle(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsStruct(const mxArray* a);
int mxIsEmpty(const mxArray* a);
int mxIsComplex(const mxArray* a);
mxArray* mxGetField(const mxArray* a, mwSize i, const char* name);
int mxGetString(const mxArray* a, char* buf, mwSize len);
void mxDestroyArray(mxArray* a);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
