// TEST INFRASTRUCTURE ONLY.  Link-time shims so that the reference sources
// compiled by ref_build.mk link without Tcl, LAPACK or BLAS (no Fortran
// compiler in this image; OTHER/LAPACK and OTHER/BLAS are .f).  None of these
// is on the measured path: Matrix.cpp keeps MATRIX_BLAS undefined
// (Matrix.cpp:52) so element products are its own loops; LAPACK is only
// reached from Matrix::Solve/Invert, which Brick/FourNodeQuad never call.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

// logging/logging.cpp needs <tcl.h>; these are the stream globals it would define
// (logging.cpp:11-17): errors to stderr, debug/log to a sink.
#include <StandardStream.h>
#include <DummyStream.h>
static StandardStream sserr;
static DummyStream    ssnul;
OPS_Stream *opserrPtr = &sserr;
OPS_Stream *opsdbgPtr = &ssnul;
OPS_Stream *opslogPtr = &ssnul;
OPS_Stream *opswrnPtr = &sserr;
OPS_Stream *opsmrdPtr = &sserr;

extern "C" {

// ---- LAPACK / BLAS subset (column-major, textbook partial pivoting) ----
void dgetrf_(int* M, int* N, double* A, int* LDA, int* ipiv, int* info) {
  int m = *M, n = *N, lda = *LDA; *info = 0;
  int mn = m < n ? m : n;
  for (int c = 0; c < mn; c++) {
    int p = c; double best = std::fabs(A[c + (size_t)c * lda]);
    for (int r = c + 1; r < m; r++)
      if (std::fabs(A[r + (size_t)c * lda]) > best) { best = std::fabs(A[r + (size_t)c * lda]); p = r; }
    ipiv[c] = p + 1;
    if (best == 0.0) { if (*info == 0) *info = c + 1; continue; }
    if (p != c) for (int k = 0; k < n; k++) { double t = A[c + (size_t)k * lda]; A[c + (size_t)k * lda] = A[p + (size_t)k * lda]; A[p + (size_t)k * lda] = t; }
    for (int r = c + 1; r < m; r++) {
      double f = (A[r + (size_t)c * lda] /= A[c + (size_t)c * lda]);
      for (int k = c + 1; k < n; k++) A[r + (size_t)k * lda] -= f * A[c + (size_t)k * lda];
    }
  }
}

int dgetrs_(char* TRANS, int* N, int* NRHS, double* A, int* LDA, int* ipiv, double* B, int* LDB, int* info) {
  int n = *N, nrhs = *NRHS, lda = *LDA, ldb = *LDB; *info = 0;
  for (int j = 0; j < nrhs; j++) {
    double* b = B + (size_t)j * ldb;
    for (int c = 0; c < n; c++) { int p = ipiv[c] - 1; if (p != c) { double t = b[c]; b[c] = b[p]; b[p] = t; } }
    for (int c = 0; c < n; c++) for (int r = c + 1; r < n; r++) b[r] -= A[r + (size_t)c * lda] * b[c];
    for (int c = n - 1; c >= 0; c--) { b[c] /= A[c + (size_t)c * lda]; for (int r = 0; r < c; r++) b[r] -= A[r + (size_t)c * lda] * b[c]; }
  }
  return 0;
}

int dgesv_(int* N, int* NRHS, double* A, int* LDA, int* ipiv, double* B, int* LDB, int* info) {
  dgetrf_(N, N, A, LDA, ipiv, info);
  if (*info != 0) return 0;
  char t = 'N';
  dgetrs_(&t, N, NRHS, A, LDA, ipiv, B, LDB, info);
  return 0;
}

void dgetri_(int* N, double* A, int* LDA, int* ipiv, double* work, int* lwork, int* info) {
  int n = *N, lda = *LDA; *info = 0;
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) inv[i + (size_t)i * n] = 1.0;
  int nrhs = n; char t = 'N';
  dgetrs_(&t, N, &nrhs, A, LDA, ipiv, inv.data(), &n, info);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[i + (size_t)j * lda] = inv[i + (size_t)j * n];
}

void dgemm_(const char* ta, const char* tb, int* M, int* N, int* K, double* alpha, double* A, const int* lda,
            double* B, const int* ldb, double* beta, double* C, const int* ldc) {
  bool tA = (*ta == 'T' || *ta == 't'), tB = (*tb == 'T' || *tb == 't');
  for (int j = 0; j < *N; j++)
    for (int i = 0; i < *M; i++) {
      double s = 0.0;
      for (int k = 0; k < *K; k++) {
        double a = tA ? A[k + (size_t)i * *lda] : A[i + (size_t)k * *lda];
        double b = tB ? B[j + (size_t)k * *ldb] : B[k + (size_t)j * *ldb];
        s += a * b;
      }
      C[i + (size_t)j * *ldc] = *alpha * s + (*beta == 0.0 ? 0.0 : *beta * C[i + (size_t)j * *ldc]);
    }
}

// ---- interpreter API the element/material "OPS_*" factory functions reference;
//      the harness constructs objects directly and never calls those factories ----
static void never(const char* what) { fprintf(stderr, "oracle/ref_shims: %s reached\n", what); abort(); }
int OPS_GetNumRemainingInputArgs() { never("OPS_GetNumRemainingInputArgs"); return 0; }
int ops_getdoubleinput_(int*, double*) { never("OPS_GetDoubleInput"); return -1; }
int ops_getintinput_(int*, int*) { never("OPS_GetIntInput"); return -1; }
const char* ops_getstring() { never("OPS_GetString"); return ""; }

// DataOutputFileHandler.cpp does not compile stand-alone (missing class tag);
// a recorder references its constructor, nothing in the harness creates one.
void shim_DataOutputFileHandler_ctor() asm("_ZN21DataOutputFileHandlerC1EPKc8echoMode8openMode");
void shim_DataOutputFileHandler_ctor() { never("DataOutputFileHandler"); }
}
