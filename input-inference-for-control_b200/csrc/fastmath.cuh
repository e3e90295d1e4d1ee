// Branch-free fp64 primitives for the inner loops.  The CUDA math library versions of rsqrt / sincos / division
// carry slow-path calls and branches (denormals, huge arguments) that break the straight-line code of the cell
// updates into many basic blocks and stop ptxas from interleaving independent dependency chains.  The arguments on
// this path are covariance pivots (normal range) and angles of O(1..1e3) rad, so the special cases are dead code.
#pragma once
#include <cuda_runtime.h>

namespace i2c {

// Polynomial / reduction constants live in the constant bank: fp64 instructions take a c[bank][offset] operand directly,
// whereas a literal costs two UMOVs (low / high word into a uniform register) at EVERY use -- 164 UMOVs = 12 % of the
// instructions of a pendulum forward cell (profiles/r01d source view).  Same values, same operation order: bit-identical.
static __constant__ double kFm[40] = {
    6.36619772367581382433e-01,   // 0  2/pi
    1.5707963267948966,           // 1  P1 = fl(pi/2)
    6.123233995736766e-17,        // 2  P2 = fl(pi/2 - P1)
    -1.4973849048591698e-33,      // 3  P3 = fl(pi/2 - P1 - P2)
    -1.66666666666666324348e-01,  // 4  S1
    8.33333333332248946124e-03,   // 5  S2
    -1.98412698298579493134e-04,  // 6  S3
    2.75573137070700676789e-06,   // 7  S4
    -2.50507602534068634195e-08,  // 8  S5
    1.58969099521155010221e-10,   // 9  S6
    4.16666666666666019037e-02,   // 10 C1
    -1.38888888888741095749e-03,  // 11 C2
    2.48015872894767294178e-05,   // 12 C3
    -2.75573143513906633035e-07,  // 13 C4
    2.08757232129817482790e-09,   // 14 C5
    -1.13596475577881948265e-11,  // 15 C6
    0.375, 0.5, 1.0, -0.5,        // 16..19
    1.0e300, 1.0e-150, 1.0e150,   // 20..22 range guards (Cholesky pivots, log-det accumulator)
    1.4426950408889634e+00,       // 23 log2(e)
    6.93147180369123816490e-01,   // 24 ln2 high part (fdlibm split)
    1.90821492927058770002e-10,   // 25 ln2 low part
    6755399441055744.0,           // 26 2^52 + 2^51: round-to-nearest-integer magic number
    -708.0,                       // 27 clamp: exp(-708) = 3e-308 is still normal
    5.00000000000000000e-01,
    1.66666666666666657e-01,
    4.16666666666666644e-02,
    8.33333333333333322e-03,
    1.38888888888888894e-03,
    1.98412698412698413e-04,
    2.48015873015873016e-05,
    2.75573192239858925e-06,
    2.75573192239858883e-07,
    2.50521083854417202e-08,
    2.08767569878681002e-09,
    1.60590438368216133e-10};  // 28..39: 1/2! .. 1/13!

// 1/sqrt(d) for normal positive d: MUFU.RSQ64H seed (~2^-22) + one third-order step (error ~ e^3 < 2^-66) => <= 1-2 ulp.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = d * y;
  const double e = fma(-h, y, kFm[18]);
  const double p = fma(kFm[16], e, kFm[17]);
  return fma(y * e, p, y);
}

// 1/d for normal d: MUFU.RCP64H seed + two Newton steps.
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, kFm[18]);
  y = fma(y, e, y);
  e = fma(-d, y, kFm[18]);
  return fma(y, e, y);
}

// exp(x) for x <= 0 (the pdf ratio exp(-q/2) of the feedback prior): round-to-nearest reduction x = k ln2 + r, |r| <= 0.347,
// degree-13 Taylor polynomial (truncation 4e-18) and
// the scaling by 2^k added straight into the exponent field.  Branch-free; arguments below -708 are clamped (3e-308, i.e. 0
// for every use here; the library routine spent 58 % of its calls in a divergent underflow path).  <= 2 ulp.
__device__ __forceinline__ double fast_exp_neg(double x) {
  x = fmax(x, kFm[27]);
  const double km = fma(x, kFm[23], kFm[26]);   // k + 2^52 + 2^51 (k in the low word)
  const double kd = km - kFm[26];
  double r = fma(-kd, kFm[24], x);
  r = fma(-kd, kFm[25], r);
  // Horner: used by the throughput variants only, where registers matter more than the length of this chain
  double t = kFm[39];
#pragma unroll
  for (int i = 38; i >= 28; --i) t = fma(t, r, kFm[i]);
  const double pe = fma(r * r, t, r) + kFm[18];
  const int k = __double2loint(km);
  return __hiloint2double(__double2hiint(pe) + (k << 20), __double2loint(pe));
}

// Latency-regime flavour of the same function: Estrin evaluation (dependency depth 7 instead of 13 fused multiply-adds) and
// no clamp on the critical path -- fp64 max() is a four-instruction sequence on this machine; arguments below -708 are
// flushed to 0 by a select on the final result, whose compare runs beside the chain.  Same polynomial: results agree with
// fast_exp_neg to <= 1 ulp (different summation order).
__device__ __forceinline__ double fast_exp_neg_lat(double x) {
  const double km = fma(x, kFm[23], kFm[26]);
  const double kd = km - kFm[26];
  double r = fma(-kd, kFm[24], x);
  r = fma(-kd, kFm[25], r);
  const double r2 = r * r;
  const double a0 = fma(kFm[29], r, kFm[28]), a1 = fma(kFm[31], r, kFm[30]), a2 = fma(kFm[33], r, kFm[32]);
  const double a3 = fma(kFm[35], r, kFm[34]), a4 = fma(kFm[37], r, kFm[36]), a5 = fma(kFm[39], r, kFm[38]);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
  const double t = fma(fma(b2, r4, b1), r4, b0);
  const double pe = fma(r2, t, r) + kFm[18];
  const int k = __double2loint(km);
  const double v = __hiloint2double(__double2hiint(pe) + (k << 20), __double2loint(pe));
  return x < kFm[27] ? 0.0 : v;
}

// sin and cos, branch-free for every finite argument: round-to-nearest multiple of pi/2 by the magic-number trick (the
// quadrant is the low word of the biased sum), then a three-term Cody-Waite reduction with FULL-precision constants.  With
// FMA the first step r1 = x - k P1 is exact for |k| < 2^51 (x and k P1 are multiples of 2^-52 and |r1| < 1), so the
// classical 33-bit split of fdlibm -- and with it the |x| < 1e5 range limit and the library fall-back that used to sit
// behind a branch in every call -- is not needed: <= 1.5 ulp for |x| <= 1e12 rad (tools/micro/sincos_model.c on the host,
// tests/test_gpu_fastmath.py on the device).  The branch mattered: with it every sincos of a cell was its own basic block
// (BSSY / BRA / BSYNC), ptxas could not interleave the three independent evaluations of a sigma-point transform, and the
// jump over the slow-path code cost instruction-cache misses (profiles/r01i source view: 3 x 160-230 cycles per transform,
// serialised).  fdlibm kernel polynomials on [-pi/4, pi/4] (Estrin pairing for ILP), branch-free quadrant selection.
__device__ __forceinline__ void fast_sincos(double x, double* sp, double* cp) {
  const double P1 = kFm[1], P2 = kFm[2], P3 = kFm[3];
  const double km = fma(x, kFm[0], kFm[26]);  // k + 2^52 + 2^51
  const double kd = km - kFm[26];
  const int k = __double2loint(km);
  double r = fma(-kd, P1, x);
  r = fma(-kd, P2, r);
  r = fma(-kd, P3, r);
  const double z = r * r;
  // sin(r) = r + r z (S1 + z S2 + z^2 S3 + ...)
  const double S1 = kFm[4], S2 = kFm[5], S3 = kFm[6], S4 = kFm[7], S5 = kFm[8], S6 = kFm[9];
  const double C1 = kFm[10], C2 = kFm[11], C3 = kFm[12], C4 = kFm[13], C5 = kFm[14], C6 = kFm[15];
  const double z2 = z * z;
  // Estrin-style pairing: (S1 + z S2) + z2 ((S3 + z S4) + z2 (S5 + z S6))
  const double s01 = fma(z, S2, S1), s23 = fma(z, S4, S3), s45 = fma(z, S6, S5);
  const double ps = fma(z2, fma(z2, s45, s23), s01);
  const double c01 = fma(z, C2, C1), c23 = fma(z, C4, C3), c45 = fma(z, C6, C5);
  const double pc = fma(z2, fma(z2, c45, c23), c01);
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z, fma(z, pc, kFm[19]), kFm[18]);
  // quadrant: k mod 4 = 0: (s,c) 1: (c,-s) 2: (-s,-c) 3: (-c, s)
  const bool swap = k & 1;
  double s = swap ? cr : sr;
  double c = swap ? sr : cr;
  s = (k & 2) ? -s : s;
  c = ((k + 1) & 2) ? -c : c;
  *sp = s;
  *cp = c;
}

// The same routine behind a (never taken) range test: used by the throughput variants.  The branch makes every evaluation
// its own basic block, i.e. ptxas cannot interleave the evaluations of a transform -- which is exactly what those variants
// need: interleaving costs registers (spills at the 128 / 168-register caps: -7 % at 65 536 problems), and with 12-16 warps
// per SM other warps fill the issue slots anyway.  Arguments beyond 1e12 rad take the library routine.
__device__ __forceinline__ void seq_sincos(double x, double* sp, double* cp) {
  if (fabs(x) > 1.0e12) {
    sincos(x, sp, cp);
    return;
  }
  fast_sincos(x, sp, cp);
}

}  // namespace i2c
