// Small dense fp64 linear algebra held entirely in registers (one problem per thread).
// Symmetric matrices are stored as packed lower triangles, row-major: (i,j), i>=j -> i(i+1)/2 + j.
// Every loop has compile-time bounds and is fully unrolled so that the arrays stay in registers.
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"

namespace i2c {

__host__ __device__ constexpr int TRI(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int tix(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
__host__ __device__ constexpr int six(int i, int j) { return i >= j ? tix(i, j) : tix(j, i); }

// Cholesky factorisation rows [START, N) of a packed lower matrix, in place (rows < START already hold L
// and invd[] their reciprocal pivots).  Returns false when a pivot is not strictly positive (or NaN) --
// the LAPACK potrf failure the reference turns into LinAlgError (inference/quadrature.py:17-24).
template <int N, int START = 0>
__device__ __forceinline__ bool chol_rows(double* A, double* invd) {
  bool ok = true;
#pragma unroll
  for (int i = START; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double s = A[tix(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-A[tix(i, k)], A[tix(j, k)], s);
      A[tix(i, j)] = s * invd[j];
    }
    double d = A[tix(i, i)];
#pragma unroll
    for (int k = 0; k < i; ++k) d = fma(-A[tix(i, k)], A[tix(i, k)], d);
    ok = ok && (d > 0.0) && (d < kFm[20]);
    double r = fast_rsqrt(d);
    invd[i] = r;
    A[tix(i, i)] = d * r;
  }
  return ok;
}

// Rows [START, N) whose first START columns have ALREADY been eliminated against the finished rows (the caller did the
// forward substitution): only the trailing block is factorised.
template <int N, int START>
__device__ __forceinline__ bool chol_rows_pre(double* A, double* invd) {
  bool ok = true;
#pragma unroll
  for (int i = START; i < N; ++i) {
#pragma unroll
    for (int j = START; j < i; ++j) {
      double s = A[tix(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-A[tix(i, k)], A[tix(j, k)], s);
      A[tix(i, j)] = s * invd[j];
    }
    double d = A[tix(i, i)];
#pragma unroll
    for (int k = 0; k < i; ++k) d = fma(-A[tix(i, k)], A[tix(i, k)], d);
    ok = ok && (d > 0.0) && (d < kFm[20]);
    double r = fast_rsqrt(d);
    invd[i] = r;
    A[tix(i, i)] = d * r;
  }
  return ok;
}

// y <- L^{-1} y  (forward substitution), L packed lower with reciprocal pivots invd.
template <int N>
__device__ __forceinline__ void fwd_subst(const double* L, const double* invd, double* y) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = y[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-L[tix(i, k)], y[k], s);
    y[i] = s * invd[i];
  }
}

// y <- L^{-T} y  (backward substitution).
template <int N>
__device__ __forceinline__ void bwd_subst(const double* L, const double* invd, double* y) {
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s = fma(-L[tix(k, i)], y[k], s);
    y[i] = s * invd[i];
  }
}

// In-place inverse of a small dense N x N matrix (row-major, not necessarily symmetric) by Gauss-Jordan
// elimination with partial pivoting (np.linalg.inv / gesv in the reference's Linearize + Riccati code).
template <int N>
__device__ __forceinline__ void inv_gj(double* A) {
  double I[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) I[i] = ((i / N) == (i % N)) ? 1.0 : 0.0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    // pivot search (select, no dynamic indexing: swap rows conditionally)
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const bool sw = fabs(A[r * N + c]) > fabs(A[c * N + c]);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double t0 = A[c * N + k], t1 = A[r * N + k];
        A[c * N + k] = sw ? t1 : t0;
        A[r * N + k] = sw ? t0 : t1;
        double u0 = I[c * N + k], u1 = I[r * N + k];
        I[c * N + k] = sw ? u1 : u0;
        I[r * N + k] = sw ? u0 : u1;
      }
    }
    const double ip = 1.0 / A[c * N + c];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      A[c * N + k] *= ip;
      I[c * N + k] *= ip;
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = A[r * N + c];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        A[r * N + k] = fma(-f, A[c * N + k], A[r * N + k]);
        I[r * N + k] = fma(-f, I[c * N + k], I[r * N + k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N * N; ++i) A[i] = I[i];
}

// C = A B for small row-major matrices (R x K)(K x Cc)
template <int R, int K, int Cc>
__device__ __forceinline__ void mm(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < Cc; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i * K + k], B[k * Cc + j], s);
      C[i * Cc + j] = s;
    }
}

// Running log-determinant accumulator that avoids one fp64 log() per pivot: keeps a renormalised
// mantissa product and an integer exponent; value() = log(product).
struct LogAcc {
  double m;
  int e;
  __device__ __forceinline__ void reset() { m = 1.0; e = 0; }
  __device__ __forceinline__ void mul(double x) {
    // branch-free renormalisation: move the exponent of the running product into e after every factor (five integer
    // instructions; a data-dependent branch here split the cell updates into extra basic blocks).  Factors are Cholesky
    // pivots: positive and normal whenever the problem's status is OK.
    m *= x;
    const int hi = __double2hiint(m);
    e += ((hi >> 20) & 0x7ff) - 1023;
    m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
  }
  __device__ __forceinline__ double value() const { return log(m) + 0.6931471805599453 * (double)e; }
};

}  // namespace i2c
