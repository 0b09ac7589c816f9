/* STUB of MATLAB's mex.h — declarations only, just enough to syntax/type-check
 * matlab/gnsscorr_mex.c in an image without MATLAB (gcc -fsyntax-only).  Never link against it. */
#ifndef GC_STUB_MEX_H
#define GC_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT8_CLASS = 8, mxUINT8_CLASS = 9, mxINT32_CLASS = 12 } mxClassID;
mxArray* mxGetField(const mxArray*, mwSize, const char*);
void mxSetField(mxArray*, mwSize, const char*, mxArray*);
double mxGetScalar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mwSize mxGetNumberOfElements(const mxArray*);
mwSize mxGetM(const mxArray*);
mwSize mxGetN(const mxArray*);
int mxIsDouble(const mxArray*);
double* mxGetDoubles(const mxArray*);
int8_t* mxGetInt8s(const mxArray*);
int32_t* mxGetInt32s(const mxArray*);
int mxIsInt8(const mxArray*);
int mxIsUint8(const mxArray*);
int mxIsChar(const mxArray*);
int mxIsStruct(const mxArray*);
size_t mxGetElementSize(const mxArray*);
int mexAtExit(void (*fn)(void));
int mxIsInt16(const mxArray*);
void* mxGetData(const mxArray*);
int mxIsEmpty(const mxArray*);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
#endif
