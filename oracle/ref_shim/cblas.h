// Shim cblas.h: the standard CBLAS prototypes the reference calls
// (src/caffe/util/math_functions.cpp:88-260 via include/caffe/util/mkl_alternate.hpp:14).
// Implemented in cblas_shim.cpp: gemm/gemv forward to an OpenBLAS found at run time
// (dlopen), otherwise to a bundled blocked OpenMP implementation.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef enum CBLAS_ORDER CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE;
void cblas_sgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE,
                 const int M, const int N, const int K, const float alpha, const float* A,
                 const int lda, const float* B, const int ldb, const float beta, float* C,
                 const int ldc);
void cblas_dgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE,
                 const int M, const int N, const int K, const double alpha, const double* A,
                 const int lda, const double* B, const int ldb, const double beta, double* C,
                 const int ldc);
void cblas_sgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int M, const int N,
                 const float alpha, const float* A, const int lda, const float* X, const int incX,
                 const float beta, float* Y, const int incY);
void cblas_dgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int M, const int N,
                 const double alpha, const double* A, const int lda, const double* X,
                 const int incX, const double beta, double* Y, const int incY);
void cblas_saxpy(const int N, const float alpha, const float* X, const int incX, float* Y,
                 const int incY);
void cblas_daxpy(const int N, const double alpha, const double* X, const int incX, double* Y,
                 const int incY);
void cblas_sscal(const int N, const float alpha, float* X, const int incX);
void cblas_dscal(const int N, const double alpha, double* X, const int incX);
float cblas_sdot(const int N, const float* X, const int incX, const float* Y, const int incY);
double cblas_ddot(const int N, const double* X, const int incX, const double* Y, const int incY);
float cblas_sasum(const int N, const float* X, const int incX);
double cblas_dasum(const int N, const double* X, const int incX);
void cblas_scopy(const int N, const float* X, const int incX, float* Y, const int incY);
void cblas_dcopy(const int N, const double* X, const int incX, double* Y, const int incY);
// which backend serves gemm: "openblas:<path>" or "builtin"
const char* mscnn_ref_blas_backend(void);
int mscnn_ref_blas_threads(void);
#ifdef __cplusplus
}
#endif
