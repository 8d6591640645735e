// CBLAS provider for the verbatim reference build (see cblas.h in this directory).
#include "cblas.h"

#include <dlfcn.h>
#include <glob.h>
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float*, int, const float*, int,
                         float, float*, int);
typedef void (*dgemm_fn)(int, int, int, int, int, int, double, const double*, int, const double*,
                         int, double, double*, int);
typedef void (*sgemv_fn)(int, int, int, int, float, const float*, int, const float*, int, float,
                         float*, int);
typedef void (*dgemv_fn)(int, int, int, int, double, const double*, int, const double*, int,
                         double, double*, int);

struct Backend {
  sgemm_fn sgemm = nullptr;
  dgemm_fn dgemm = nullptr;
  sgemv_fn sgemv = nullptr;
  dgemv_fn dgemv = nullptr;
  std::string name = "builtin";
  int threads = 1;
  void (*set_threads)(int) = nullptr;
  int (*get_threads)(void) = nullptr;
};

Backend load_backend() {
  Backend b;
  b.threads = omp_get_max_threads();
  if (std::getenv("MSCNN_REF_BLAS_BUILTIN")) return b;
  std::vector<std::string> cands;
  if (const char* e = std::getenv("MSCNN_REF_BLAS")) cands.push_back(e);
  const char* pats[] = {
      // LP64 OpenBLAS bundled with scipy (symbols prefixed scipy_), then the one bundled with
      // opencv (plain cblas_ symbols), then a system install.
      "/opt/prime-rl/.venv/lib/python3*/site-packages/scipy.libs/libscipy_openblas-*.so",
      "/opt/prime-rl/.venv/lib/python3*/site-packages/opencv_python_headless.libs/libopenblas*.so*",
      "/usr/lib/x86_64-linux-gnu/libopenblas.so*", "/usr/lib/x86_64-linux-gnu/openblas*/libopenblas*.so*"};
  for (const char* p : pats) {
    glob_t g;
    if (glob(p, 0, nullptr, &g) == 0) {
      for (size_t i = 0; i < g.gl_pathc; ++i) cands.push_back(g.gl_pathv[i]);
    }
    globfree(&g);
  }
  for (const std::string& c : cands) {
    // wheels bundle their Fortran runtime next to the library without an RPATH: load it first
    const std::string dir = c.substr(0, c.find_last_of('/') + 1);
    for (const char* dep : {"libquadmath*.so*", "libgfortran*.so*"}) {
      glob_t g;
      if (glob((dir + dep).c_str(), 0, nullptr, &g) == 0)
        for (size_t i = 0; i < g.gl_pathc; ++i) dlopen(g.gl_pathv[i], RTLD_NOW | RTLD_GLOBAL);
      globfree(&g);
    }
    void* h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) continue;
    for (const char* prefix : {"", "scipy_"}) {
      const std::string pf = prefix;
      sgemm_fn s = reinterpret_cast<sgemm_fn>(dlsym(h, (pf + "cblas_sgemm").c_str()));
      dgemm_fn d = reinterpret_cast<dgemm_fn>(dlsym(h, (pf + "cblas_dgemm").c_str()));
      sgemv_fn sv = reinterpret_cast<sgemv_fn>(dlsym(h, (pf + "cblas_sgemv").c_str()));
      dgemv_fn dv = reinterpret_cast<dgemv_fn>(dlsym(h, (pf + "cblas_dgemv").c_str()));
      if (s && d && sv && dv) {
        b.sgemm = s; b.dgemm = d; b.sgemv = sv; b.dgemv = dv;
        b.name = "openblas:" + c;
        typedef int (*nthr_fn)(void);
        if (nthr_fn nt = reinterpret_cast<nthr_fn>(dlsym(h, (pf + "openblas_get_num_threads").c_str()))) {
          b.threads = nt();
          b.get_threads = nt;
        }
        b.set_threads = reinterpret_cast<void (*)(int)>(dlsym(h, (pf + "openblas_set_num_threads").c_str()));
        return b;
      }
    }
    dlclose(h);
  }
  return b;
}

Backend& backend() {
  static Backend b = load_backend();
  return b;
}

// C(MxN) = alpha * op(A) * op(B) + beta * C, row major.  Blocked, OpenMP over row panels.
template <typename T>
void gemm_builtin(bool ta, bool tb, int M, int N, int K, T alpha, const T* A, int lda, const T* B,
                  int ldb, T beta, T* C, int ldc) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    T* c = C + (size_t)i * ldc;
    if (beta == T(0)) std::fill(c, c + N, T(0));
    else if (beta != T(1)) for (int j = 0; j < N; ++j) c[j] *= beta;
  }
  const int MB = 48, KB = 256;
#pragma omp parallel for schedule(dynamic)
  for (int i0 = 0; i0 < M; i0 += MB) {
    const int i1 = std::min(M, i0 + MB);
    std::vector<T> brow;
    for (int k0 = 0; k0 < K; k0 += KB) {
      const int k1 = std::min(K, k0 + KB);
      for (int i = i0; i < i1; ++i) {
        T* c = C + (size_t)i * ldc;
        for (int k = k0; k < k1; ++k) {
          const T a = alpha * (ta ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]);
          if (!tb) {
            const T* b = B + (size_t)k * ldb;
            for (int j = 0; j < N; ++j) c[j] += a * b[j];
          } else {
            for (int j = 0; j < N; ++j) c[j] += a * B[(size_t)j * ldb + k];
          }
        }
      }
    }
  }
}

template <typename T>
void gemv_builtin(bool ta, int M, int N, T alpha, const T* A, int lda, const T* x, T beta, T* y) {
  const int ylen = ta ? N : M;
  for (int i = 0; i < ylen; ++i) y[i] = (beta == T(0)) ? T(0) : beta * y[i];
  if (!ta) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
      T s = 0;
      for (int j = 0; j < N; ++j) s += A[(size_t)i * lda + j] * x[j];
      y[i] += alpha * s;
    }
  } else {
    for (int i = 0; i < M; ++i) {
      const T a = alpha * x[i];
      for (int j = 0; j < N; ++j) y[j] += a * A[(size_t)i * lda + j];
    }
  }
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* mscnn_ref_blas_backend(void) { return backend().name.c_str(); }
int mscnn_ref_blas_threads(void) { return backend().threads; }
// Explicit thread count for the BLAS behind the reference layers (torchrun exports OMP_NUM_THREADS=1, which would
// silently turn the "all host cores" baseline into a single-threaded one).  Returns the count in effect.
int mscnn_ref_blas_set_threads(int n) {
  Backend& b = backend();
  if (n < 1) n = 1;
  if (b.set_threads) {
    b.set_threads(n);
    b.threads = b.get_threads ? b.get_threads() : n;
  } else {
    omp_set_num_threads(n);
    b.threads = omp_get_max_threads();
  }
  return b.threads;
}

void cblas_sgemm(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const float alpha, const float* A,
                 const int lda, const float* B, const int ldb, const float beta, float* C,
                 const int ldc) {
  if (backend().sgemm) { backend().sgemm(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); return; }
  if (o != CblasRowMajor) { fprintf(stderr, "cblas shim: row-major only\n"); abort(); }
  gemm_builtin<float>(ta != CblasNoTrans, tb != CblasNoTrans, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
void cblas_dgemm(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const double alpha, const double* A,
                 const int lda, const double* B, const int ldb, const double beta, double* C,
                 const int ldc) {
  if (backend().dgemm) { backend().dgemm(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); return; }
  if (o != CblasRowMajor) { fprintf(stderr, "cblas shim: row-major only\n"); abort(); }
  gemm_builtin<double>(ta != CblasNoTrans, tb != CblasNoTrans, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
void cblas_sgemv(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE ta, const int M, const int N,
                 const float alpha, const float* A, const int lda, const float* X, const int incX,
                 const float beta, float* Y, const int incY) {
  if (backend().sgemv) { backend().sgemv(o, ta, M, N, alpha, A, lda, X, incX, beta, Y, incY); return; }
  if (o != CblasRowMajor || incX != 1 || incY != 1) { fprintf(stderr, "cblas shim: gemv form\n"); abort(); }
  gemv_builtin<float>(ta != CblasNoTrans, M, N, alpha, A, lda, X, beta, Y);
}
void cblas_dgemv(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE ta, const int M, const int N,
                 const double alpha, const double* A, const int lda, const double* X,
                 const int incX, const double beta, double* Y, const int incY) {
  if (backend().dgemv) { backend().dgemv(o, ta, M, N, alpha, A, lda, X, incX, beta, Y, incY); return; }
  if (o != CblasRowMajor || incX != 1 || incY != 1) { fprintf(stderr, "cblas shim: gemv form\n"); abort(); }
  gemv_builtin<double>(ta != CblasNoTrans, M, N, alpha, A, lda, X, beta, Y);
}
void cblas_saxpy(const int N, const float a, const float* X, const int ix, float* Y, const int iy) {
  for (int i = 0; i < N; ++i) Y[(size_t)i * iy] += a * X[(size_t)i * ix];
}
void cblas_daxpy(const int N, const double a, const double* X, const int ix, double* Y, const int iy) {
  for (int i = 0; i < N; ++i) Y[(size_t)i * iy] += a * X[(size_t)i * ix];
}
void cblas_sscal(const int N, const float a, float* X, const int ix) {
  for (int i = 0; i < N; ++i) X[(size_t)i * ix] *= a;
}
void cblas_dscal(const int N, const double a, double* X, const int ix) {
  for (int i = 0; i < N; ++i) X[(size_t)i * ix] *= a;
}
float cblas_sdot(const int N, const float* X, const int ix, const float* Y, const int iy) {
  float s = 0;
  for (int i = 0; i < N; ++i) s += X[(size_t)i * ix] * Y[(size_t)i * iy];
  return s;
}
double cblas_ddot(const int N, const double* X, const int ix, const double* Y, const int iy) {
  double s = 0;
  for (int i = 0; i < N; ++i) s += X[(size_t)i * ix] * Y[(size_t)i * iy];
  return s;
}
float cblas_sasum(const int N, const float* X, const int ix) {
  float s = 0;
  for (int i = 0; i < N; ++i) s += X[(size_t)i * ix] < 0 ? -X[(size_t)i * ix] : X[(size_t)i * ix];
  return s;
}
double cblas_dasum(const int N, const double* X, const int ix) {
  double s = 0;
  for (int i = 0; i < N; ++i) s += X[(size_t)i * ix] < 0 ? -X[(size_t)i * ix] : X[(size_t)i * ix];
  return s;
}
void cblas_scopy(const int N, const float* X, const int ix, float* Y, const int iy) {
  for (int i = 0; i < N; ++i) Y[(size_t)i * iy] = X[(size_t)i * ix];
}
void cblas_dcopy(const int N, const double* X, const int ix, double* Y, const int iy) {
  for (int i = 0; i < N; ++i) Y[(size_t)i * iy] = X[(size_t)i * ix];
}
}
#pragma GCC visibility pop
