// Branch-free fp64 primitives for the inner loops.  The CUDA math library versions of rsqrt / sincos / division
// carry slow-path calls and branches (denormals, huge arguments) that break the straight-line code of the cell
// updates into many basic blocks and stop ptxas from interleaving independent dependency chains.  The arguments on
// this path are covariance pivots (normal range) and angles of O(1..1e3) rad, so the special cases are dead code.
#pragma once
#include <cuda_runtime.h>

namespace i2c {

// 1/sqrt(d) for normal positive d: MUFU.RSQ64H seed (~2^-22) + one third-order step (error ~ e^3 < 2^-66) => <= 1-2 ulp.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = d * y;
  const double e = fma(-h, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// 1/d for normal d: MUFU.RCP64H seed + two Newton steps.
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
}

// sin and cos for |x| < ~1e5: three-term Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials on
// [-pi/4, pi/4] (Horner split in even/odd halves for ILP).  Branch-free quadrant selection.  Larger arguments fall
// back to the library routine (warp-uniform in practice: never taken for the registered environments).
__device__ __forceinline__ void fast_sincos(double x, double* sp, double* cp) {
  if (fabs(x) > 1.0e5) {
    sincos(x, sp, cp);
    return;
  }
  const double TWO_OVER_PI = 6.36619772367581382433e-01;
  const double P1 = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double P2 = 6.07710050650619224932e-11;  // pi/2 - P1, first 33 bits
  const double P3 = 2.02226624879595063154e-21;  // pi/2 - (P1 + P2)
  const double kd = rint(x * TWO_OVER_PI);
  const int k = (int)kd;
  double r = fma(-kd, P1, x);
  r = fma(-kd, P2, r);
  r = fma(-kd, P3, r);
  const double z = r * r;
  // sin(r) = r + r z (S1 + z S2 + z^2 S3 + ...)
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z2 = z * z;
  // Estrin-style pairing: (S1 + z S2) + z2 ((S3 + z S4) + z2 (S5 + z S6))
  const double s01 = fma(z, S2, S1), s23 = fma(z, S4, S3), s45 = fma(z, S6, S5);
  const double ps = fma(z2, fma(z2, s45, s23), s01);
  const double c01 = fma(z, C2, C1), c23 = fma(z, C4, C3), c45 = fma(z, C6, C5);
  const double pc = fma(z2, fma(z2, c45, c23), c01);
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z, fma(z, pc, -0.5), 1.0);
  // quadrant: k mod 4 = 0: (s,c) 1: (c,-s) 2: (-s,-c) 3: (-c, s)
  const bool swap = k & 1;
  double s = swap ? cr : sr;
  double c = swap ? sr : cr;
  s = (k & 2) ? -s : s;
  c = ((k + 1) & 2) ? -c : c;
  *sp = s;
  *cp = c;
}

}  // namespace i2c
