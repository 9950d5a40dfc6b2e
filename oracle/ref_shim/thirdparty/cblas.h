/* Declarations of the CBLAS entry points the reference calls (math_functions.cpp, mkl_alternate.hpp).
 * TEST INFRASTRUCTURE for oracle/_ref: no CBLAS headers exist in the image; the definitions in ref_capi.cpp forward to
 * the OpenBLAS that ships inside the image's numpy/scipy wheels (dlopen at run time), i.e. a real `BLAS := open`. */
#ifndef REF_SHIM_CBLAS_H_
#define REF_SHIM_CBLAS_H_
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, float alpha, const float* A, int lda,
                 const float* B, int ldb, float beta, float* C, int ldc);
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, double alpha, const double* A, int lda,
                 const double* B, int ldb, double beta, double* C, int ldc);
void cblas_sgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int M, int N, float alpha, const float* A, int lda, const float* x, int incx,
                 float beta, float* y, int incy);
void cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int M, int N, double alpha, const double* A, int lda, const double* x, int incx,
                 double beta, double* y, int incy);
void cblas_saxpy(int n, float a, const float* x, int incx, float* y, int incy);
void cblas_daxpy(int n, double a, const double* x, int incx, double* y, int incy);
void cblas_sscal(int n, float a, float* x, int incx);
void cblas_dscal(int n, double a, double* x, int incx);
void cblas_scopy(int n, const float* x, int incx, float* y, int incy);
void cblas_dcopy(int n, const double* x, int incx, double* y, int incy);
float cblas_sdot(int n, const float* x, int incx, const float* y, int incy);
double cblas_ddot(int n, const double* x, int incx, const double* y, int incy);
float cblas_sasum(int n, const float* x, int incx);
double cblas_dasum(int n, const double* x, int incx);
#endif
