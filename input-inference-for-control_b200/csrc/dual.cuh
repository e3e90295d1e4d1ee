// Forward-mode automatic differentiation with a fixed number of partials: the Linearize inference of the reference
// takes Jacobians of the environment dynamics with autograd (i2c/env_autograd.py:22,57,170; i2c/model.py:158-164);
// here the same dynamics are evaluated once on dual numbers inside the kernel.
#pragma once
#include <cuda_runtime.h>

namespace i2c {

template <int NV>
struct Dual {
  double v;
  double d[NV];
};

template <int NV>
__device__ __forceinline__ Dual<NV> dconst(double c) {
  Dual<NV> r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = 0.0;
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> dvar(double x, int idx) {
  Dual<NV> r = dconst<NV>(x);
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = (i == idx) ? 1.0 : 0.0;
  return r;
}
// r = f(a) with derivative df
template <int NV>
__device__ __forceinline__ Dual<NV> dchain(const Dual<NV>& a, double f, double df) {
  Dual<NV> r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = df * a.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator+(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a) { return dchain(a, -a.v, -1.0); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator*(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v);
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator/(const Dual<NV>& a, const Dual<NV>& b) {
  const double ib = 1.0 / b.v, q = a.v * ib;
  Dual<NV> r;
  r.v = q;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = (a.d[i] - q * b.d[i]) * ib;
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator+(const Dual<NV>& a, double c) { return dchain(a, a.v + c, 1.0); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator+(double c, const Dual<NV>& a) { return dchain(a, a.v + c, 1.0); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a, double c) { return dchain(a, a.v - c, 1.0); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator-(double c, const Dual<NV>& a) { return dchain(a, c - a.v, -1.0); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator*(const Dual<NV>& a, double c) { return dchain(a, a.v * c, c); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator*(double c, const Dual<NV>& a) { return dchain(a, a.v * c, c); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator/(const Dual<NV>& a, double c) { return dchain(a, a.v / c, 1.0 / c); }

// scalar-generic elementary functions (double and Dual overloads)
__device__ __forceinline__ double t_sin(double x) { return sin(x); }
__device__ __forceinline__ double t_cos(double x) { return cos(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double t_clip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
template <int NV>
__device__ __forceinline__ Dual<NV> t_sin(const Dual<NV>& a) {
  double s, c;
  sincos(a.v, &s, &c);
  return dchain(a, s, c);
}
template <int NV>
__device__ __forceinline__ Dual<NV> t_cos(const Dual<NV>& a) {
  double s, c;
  sincos(a.v, &s, &c);
  return dchain(a, c, -s);
}
template <int NV>
__device__ __forceinline__ Dual<NV> t_sqrt(const Dual<NV>& a) {
  const double r = sqrt(a.v);
  return dchain(a, r, 0.5 / r);
}
// np.clip: the gradient passes where lo <= x <= hi (autograd's rule)
template <int NV>
__device__ __forceinline__ Dual<NV> t_clip(const Dual<NV>& a, double lo, double hi) {
  return dchain(a, fmin(fmax(a.v, lo), hi), (a.v >= lo && a.v <= hi) ? 1.0 : 0.0);
}

}  // namespace i2c
