// In-kernel environment definitions: dynamics (one Euler step, input clipped), cost-feature maps
// observe / observe_terminal, and the quadrotor measurement map.  Restated from the reference's
// NumPy functions (cited per env; paths relative to the reference root).
//
// Sigma points are m +- sf*L[:,j] with L lower triangular, so coordinate i of the points built from
// column j > i equals the centre m_i exactly: the trigonometric features of an angle stored at state
// index a only need fresh sin/cos for columns j <= a.  Every map therefore takes the column index j
// (compile-time after unrolling; j < 0 = centre) and a `Trig` cache evaluated once at the centre.
#pragma once
#include "dual.cuh"
#include "linalg.cuh"

namespace i2c {

// MODE bit 0: branch-free sincos (latency variants: independent evaluations interleave) instead of the sequenced one
//              (throughput variants: lower register pressure);
// MODE bit 1: angle addition ("PM").  The sigma points of column j move an angle by +-d, d = sf L[angle][j], so
//              sin / cos(m +- d) = s_m c_d +- c_m s_d / c_m c_d -+ s_m s_d: ONE sincos of the offset serves both points of
//              the column (offsets(): once per transform).  The minus point is announced to the env maps by bit 6 of
//              the column index (kMinus).
constexpr int kMinus = 64;

// np.clip for the action limits: two compares and selects.  fmin(fmax()) carries IEEE NaN handling that costs ~8 instructions
// per call on this machine (DSETP.MAX + a select / sign-fix sequence per 32-bit half): 100 of the ~220 instructions of a
// pendulum dynamics transform.  A NaN input stays NaN, as in np.clip.
__device__ __forceinline__ double clipd(double x, double lo, double hi) {
  double y = x > hi ? hi : x;
  return x < lo ? lo : y;
}
template <int NA, int MODE = 0, int NJ = 1>
struct Trig {
  static constexpr bool kFree = MODE & 1, kPM = (MODE & 2) != 0;
  static constexpr int A = NA > 0 ? NA : 1, J = NJ > 0 ? NJ : 1;
  double s[A], c[A];
  double ds[kPM ? A * J : 1], dc[kPM ? A * J : 1];
  __device__ __forceinline__ void sc(double x, double* sp, double* cp) const {
    if constexpr (kFree) fast_sincos(x, sp, cp);
    else seq_sincos(x, sp, cp);
  }
  // sin / cos of (centre angle a) +- (offset of column j)
  __device__ __forceinline__ void pm(int a, int j, bool minus, double& so, double& co) const {
    const double sd = minus ? -ds[a * J + j] : ds[a * J + j], cd = dc[a * J + j];
    so = fma(s[a], cd, c[a] * sd);
    co = fma(c[a], cd, -(s[a] * sd));
  }
};

// ------------------------------------------------------------------ linear systems
// LinearDef / LinearBase: env_def.py:139-191, model.py:226-242.  Per-problem parameters
// par = [A00 A01 A10 A11 B0 B1 a0 a1].
struct EnvLinear {
  static constexpr int DX = 2, DU = 1, DZ = 3, DZT = 2, NP = 8, DY = 0, NA = 0;
  static constexpr bool HAS_TERM = true, LINEAR = true;
  static constexpr bool J_HANDOFF = false;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  static constexpr int NJ = 0;
  using TrigT = Trig<NA>;
  template <class TT> __device__ static void offsets(const double*, double, TT&) {}
  // z = E x + F u (+ e = 0): observe_linearize (env_def.py:171-181): E = [I2; 0], F = [0 0 1]^T
  __host__ __device__ static constexpr double obsE(int a, int i) { return a == i ? 1.0 : 0.0; }
  __host__ __device__ static constexpr double obsF(int a, int) { return a == 2 ? 1.0 : 0.0; }
  // structure of the cost-feature maps: feature a is the identity of joint-state component obs_src(a) >= 0, or the
  // (-1-k)-th nonlinear (trigonometric) feature; OBS_JMAX = last sigma-point column that changes a nonlinear feature
  static constexpr int OBS_NL = 0, OBS_JMAX = -1;
  __host__ __device__ static constexpr int obs_src(int a) { return a; }
  __host__ __device__ static constexpr int term_src(int a) { return a; }
  __host__ __device__ static constexpr int nl_angle(int) { return 0; }
  template <class T>
  __device__ static void dyn_g(const T* xu, const double* par, T* y) {
    y[0] = par[0] * xu[0] + par[1] * xu[1] + par[4] * xu[2] + par[6];
    y[1] = par[2] * xu[0] + par[3] * xu[1] + par[5] * xu[2] + par[7];
  }
  template <class TT> __device__ static void trig_nl(const double*, int, const TT&, double*) {}
  template <class TT> __device__ static void center(const double*, TT&) {}
  template <class TT> __device__ static void dyn(const double* xu, int, const TT&, const double* par, double* y) {
    y[0] = fma(par[0], xu[0], fma(par[1], xu[1], fma(par[4], xu[2], par[6])));
    y[1] = fma(par[2], xu[0], fma(par[3], xu[1], fma(par[5], xu[2], par[7])));
  }
  template <class TT> __device__ static void obs(const double* xu, int, const TT&, double* z) {
    z[0] = xu[0]; z[1] = xu[1]; z[2] = xu[2];
  }
  template <class TT> __device__ static void obs_term(const double* x, int, const TT&, double* z) { z[0] = x[0]; z[1] = x[1]; }
  template <class TT> __device__ static void measure(const double*, int, const TT&, double*) {}
};

// LinearMinimumEnergyDef: env_def.py:194-230 (cost on u only).
struct EnvLinearMinEnergy : EnvLinear {
  static constexpr int DZ = 1;
  // env_def.py:211-217: C = 0 (1x2), D = 1
  __host__ __device__ static constexpr double obsE(int, int) { return 0.0; }
  __host__ __device__ static constexpr double obsF(int, int) { return 1.0; }
  __host__ __device__ static constexpr int obs_src(int) { return 2; }
  template <class TT> __device__ static void obs(const double* xu, int, const TT&, double* z) { z[0] = xu[2]; }
};

// Literals of the hot dynamics live in the constant bank (fp64 instructions take c[bank][offset] operands directly; a
// literal costs two UMOVs at every use).  The initialisers are the same constant expressions the code used, folded by the
// same compiler in the same order => bit-identical results.
static __constant__ double kPendulum[6] = {0.05, 1e-2, -3.0 * 9.80665 / 2.0, 3.0, -2.0, 2.0};
static __constant__ double kDcp[15] = {
    1.0 / 125.0,
    0.127 * (0.3365 / 2) + 0.127 * 0.3365,                                   // a12 = Mp1 l1 + Mp2 L2
    0.127 * (0.3365 / 2),                                                    // a13 = Mp2 l2
    0.3365 * (0.3365 / 2) * 0.127,                                           // a23 = L1 l2 Mp2
    (0.3365 / 2) * (0.3365 / 2) * 0.127 + 0.3365 * 0.3365 * 0.127 + 0.127 * 0.3365 / 12,  // M22
    (0.3365 / 2) * (0.3365 / 2) * 0.127 + 0.127 * 0.3365 / 12,               // M33
    -(0.127 * (0.3365 / 2) + 0.127 * 0.3365),
    -(0.127 * (0.3365 / 2)),
    -(0.3365 * (0.3365 / 2) * 0.127),
    -(0.127 * (0.3365 / 2) + 0.127 * 0.3365) * 9.81,                         // -(Mp1 l1 + Mp2 L1) g
    -0.127 * (0.3365 / 2) * 9.81,                                            // -Mp2 l2 g
    3.0, -10.0, 10.0, 1.2659242088545832};
static __constant__ double kQuad[8] = {0.1, 1.0 / (5.0 * (2 * 0.8) * (2 * (400.0 / 30.0 / 100.0))), -9.81, 0.8,
                                       1.0 / ((5.0 * (2 * 0.8) * (2 * (400.0 / 30.0 / 100.0))) *
                                              ((2 * 0.8) * (2 * 0.8) + (2 * (400.0 / 30.0 / 100.0)) * (2 * (400.0 / 30.0 / 100.0))) / 12.0),
                                       1.0 / (1.0 + 0.1 * 0.5), 0.0, 30.0};
static __constant__ double kCartpole[10] = {-0.127 * 0.3365, (0.37 + 0.127) * 9.81, 0.3365, (4.0 / 3.0) * (0.37 + 0.127), 0.127,
                                            0.127 * 0.3365, 1.0 / (0.37 + 0.127), 1.0 / 250.0, -5.0, 5.0};

// ------------------------------------------------------------------ pendulum
// env_autograd.py:5-19 (dynamics), env_def.py:273-291 (features [sin th, cos th, thd, u]).
struct EnvPendulum {
  static constexpr int DX = 2, DU = 1, DZ = 4, DZT = 3, NP = 0, DY = 0, NA = 1;
  static constexpr bool HAS_TERM = true, LINEAR = false;
  static constexpr bool J_HANDOFF = false;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  __host__ __device__ static constexpr double obsE(int, int) { return 0.0; }
  __host__ __device__ static constexpr double obsF(int, int) { return 0.0; }
  static constexpr int NJ = 1;
  using TrigT = Trig<NA, 0, NJ>;
  template <class TT> __device__ static void center(const double* m, TT& t) { t.sc(m[0], &t.s[0], &t.c[0]); }
  template <class TT> __device__ static void offsets(const double* L, double sf, TT& t) {
    if constexpr (TT::kPM) t.sc(sf * L[tix(0, 0)], &t.ds[0], &t.dc[0]);
  }
  template <class TT> __device__ static void trig(const double* x, int j, const TT& c, double& s, double& co) {
    if (j < 0 || (j & ~kMinus) != 0) { s = c.s[0]; co = c.c[0]; }
    else if constexpr (TT::kPM) c.pm(0, 0, j & kMinus, s, co);
    else c.sc(x[0], &s, &co);
  }
  // scalar-generic dynamics (double or Dual): used by the Linearize inference to get value + Jacobian in one pass
  template <class T>
  __device__ static void dyn_g(const T* xu, const double*, T* y) {
    const double dt = 0.05, d = 1e-2, g = 9.80665;
    T u = t_clip(xu[2], -2.0, 2.0);
    T acc = (-3.0 * g / 2.0) * t_sin(xu[0] + 3.141592653589793) - d * xu[1];
    acc = acc + 3.0 * u;
    T xd = xu[1] + acc * dt;
    y[0] = xu[0] + xd * dt;
    y[1] = xd;
  }
  __host__ __device__ static constexpr int nl_angle(int) { return 0; }
  static constexpr int OBS_NL = 2, OBS_JMAX = 0;  // z = [sin th, cos th | thd, u]
  __host__ __device__ static constexpr int obs_src(int a) { return a < 2 ? -1 - a : a - 1; }
  __host__ __device__ static constexpr int term_src(int a) { return a < 2 ? -1 - a : a - 1; }
  template <class TT> __device__ static void trig_nl(const double* x, int j, const TT& c, double* y) { trig(x, j, c, y[0], y[1]); }
  template <class TT> __device__ static void dyn(const double* xu, int j, const TT& c, const double*, double* y) {
    const double dt = kPendulum[0], d = kPendulum[1];
    double s, co;
    trig(xu, j, c, s, co);
    (void)co;
    double u = clipd(xu[2], kPendulum[4], kPendulum[5]);
    // the reference evaluates np.sin(th + np.pi); sin(th + pi) == -sin(th) up to the rounding of th + pi
    double acc = kPendulum[2] * (-s) - d * xu[1];  // kPendulum[2] = -3 g / 2
    acc += kPendulum[3] * u;
    double xd = fma(acc, dt, xu[1]);
    y[0] = fma(xd, dt, xu[0]);
    y[1] = xd;
  }
  template <class TT> __device__ static void obs(const double* xu, int j, const TT& c, double* z) {
    trig(xu, j, c, z[0], z[1]);
    z[2] = xu[1];
    z[3] = xu[2];
  }
  template <class TT> __device__ static void obs_term(const double* x, int j, const TT& c, double* z) {
    trig(x, j, c, z[0], z[1]);
    z[2] = x[1];
  }
  template <class TT> __device__ static void measure(const double*, int, const TT&, double*) {}
};

// PendulumKnownActReg: env_def.py:312-346 (cost on u only, no terminal features).
struct EnvPendulumActReg : EnvPendulum {
  static constexpr int DZ = 1, DZT = 1;
  static constexpr bool HAS_TERM = false;
  static constexpr bool J_HANDOFF = false;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  static constexpr int OBS_NL = 0, OBS_JMAX = -1;
  __host__ __device__ static constexpr int obs_src(int) { return 2; }
  __host__ __device__ static constexpr int term_src(int) { return 0; }
  template <class TT> __device__ static void trig_nl(const double*, int, const TT&, double*) {}
  template <class TT> __device__ static void obs(const double* xu, int, const TT&, double* z) { z[0] = xu[2]; }
  template <class TT> __device__ static void obs_term(const double*, int, const TT&, double* z) { z[0] = 0.0; }
};

// ------------------------------------------------------------------ cart-pole
// env_autograd.py:25-54, env_def.py:537-570 (features [x, sin th, cos th, xd, thd, u]).
struct EnvCartpole {
  static constexpr int DX = 4, DU = 1, DZ = 6, DZT = 5, NP = 0, DY = 0, NA = 1;
  static constexpr bool HAS_TERM = true, LINEAR = false;
  static constexpr bool J_HANDOFF = true;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  __host__ __device__ static constexpr double obsE(int, int) { return 0.0; }
  __host__ __device__ static constexpr double obsF(int, int) { return 0.0; }
  static constexpr int NJ = 2;
  using TrigT = Trig<NA, 0, NJ>;
  template <class TT> __device__ static void center(const double* m, TT& t) { t.sc(m[1], &t.s[0], &t.c[0]); }
  template <class TT> __device__ static void offsets(const double* L, double sf, TT& t) {
    if constexpr (TT::kPM) {
#pragma unroll
      for (int j = 0; j < 2; ++j) t.sc(sf * L[tix(1, j)], &t.ds[j], &t.dc[j]);
    }
  }
  template <class TT> __device__ static void trig(const double* x, int j, const TT& c, double& s, double& co) {
    if (j < 0 || (j & ~kMinus) > 1) { s = c.s[0]; co = c.c[0]; }
    else if constexpr (TT::kPM) c.pm(0, j & ~kMinus, j & kMinus, s, co);
    else c.sc(x[1], &s, &co);
  }
  template <class T>
  __device__ static void dyn_g(const T* xu, const double*, T* y) {
    const double g = 9.81, Mc = 0.37, Mp = 0.127, Mt = Mc + Mp, l = 0.3365, dt = 1.0 / 250.0;
    T u = t_clip(xu[4], -5.0, 5.0);
    T sth = t_sin(xu[1]), cth = t_cos(xu[1]);
    T dth2 = xu[3] * xu[3];
    T num = (-Mp * l) * sth * cth * dth2 + (Mt * g) * sth - u * cth;
    T den = l * ((4.0 / 3.0) * Mt - Mp * (cth * cth));
    T th_acc = num / den;
    T x_acc = ((Mp * l) * sth * dth2 - (Mp * l) * th_acc * cth + u) / Mt;
    y[0] = xu[0] + dt * xu[2];
    y[1] = xu[1] + dt * xu[3];
    y[2] = xu[2] + dt * x_acc;
    y[3] = xu[3] + dt * th_acc;
  }
  __host__ __device__ static constexpr int nl_angle(int) { return 1; }
  static constexpr int OBS_NL = 2, OBS_JMAX = 1;  // z = [x, sin th, cos th, xd, thd, u]
  __host__ __device__ static constexpr int obs_src(int a) { return a == 0 ? 0 : (a <= 2 ? -a : a - 1); }
  __host__ __device__ static constexpr int term_src(int a) { return a == 0 ? 0 : (a <= 2 ? -a : a - 1); }
  template <class TT> __device__ static void trig_nl(const double* x, int j, const TT& c, double* y) { trig(x, j, c, y[0], y[1]); }
  template <class TT> __device__ static void dyn(const double* xu, int j, const TT& c, const double*, double* y) {
    // kCartpole = {-Mp l, Mt g, l, 4/3 Mt, Mp, Mp l, 1/Mt, dt, -5, 5}
    const double dt = kCartpole[7];
    double u = clipd(xu[4], kCartpole[8], kCartpole[9]);
    double sth, cth;
    trig(xu, j, c, sth, cth);
    double dth2 = xu[3] * xu[3];
    double num = kCartpole[0] * sth * cth * dth2 + kCartpole[1] * sth - u * cth;
    double den = kCartpole[2] * (kCartpole[3] - kCartpole[4] * cth * cth);
    double th_acc = num * fast_rcp(den);
    double x_acc = (kCartpole[5] * sth * dth2 - kCartpole[5] * th_acc * cth + u) * kCartpole[6];
    y[0] = fma(dt, xu[2], xu[0]);
    y[1] = fma(dt, xu[3], xu[1]);
    y[2] = fma(dt, x_acc, xu[2]);
    y[3] = fma(dt, th_acc, xu[3]);
  }
  template <class TT> __device__ static void obs(const double* xu, int j, const TT& c, double* z) {
    z[0] = xu[0];
    trig(xu, j, c, z[1], z[2]);
    z[3] = xu[2]; z[4] = xu[3]; z[5] = xu[4];
  }
  template <class TT> __device__ static void obs_term(const double* x, int j, const TT& c, double* z) {
    z[0] = x[0];
    trig(x, j, c, z[1], z[2]);
    z[3] = x[2]; z[4] = x[3];
  }
  template <class TT> __device__ static void measure(const double*, int, const TT&, double*) {}
};

// ------------------------------------------------------------------ double cart-pole
// env_autograd.py:60-167, env_def.py:682-732
// (features [x, sin th1, cos th1, sin th2, cos th2, xd, thd1, thd2, u]).
struct EnvDoubleCartpole {
  static constexpr int DX = 6, DU = 1, DZ = 9, DZT = 8, NP = 0, DY = 0, NA = 2;
  static constexpr bool HAS_TERM = true, LINEAR = false;
  static constexpr bool J_HANDOFF = false;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  __host__ __device__ static constexpr double obsE(int, int) { return 0.0; }
  __host__ __device__ static constexpr double obsF(int, int) { return 0.0; }
  static constexpr int NJ = 3;
  using TrigT = Trig<NA, 0, NJ>;
  template <class TT> __device__ static void center(const double* m, TT& t) {
    t.sc(m[1], &t.s[0], &t.c[0]);
    t.sc(m[2], &t.s[1], &t.c[1]);
  }
  template <class TT> __device__ static void offsets(const double* L, double sf, TT& t) {
    if constexpr (TT::kPM) {
#pragma unroll
      for (int j = 0; j < 2; ++j) t.sc(sf * L[tix(1, j)], &t.ds[j], &t.dc[j]);
#pragma unroll
      for (int j = 0; j < 3; ++j) t.sc(sf * L[tix(2, j)], &t.ds[3 + j], &t.dc[3 + j]);
    }
  }
  template <class TT> __device__ static void trig1(const double* x, int j, const TT& c, double& s, double& co) {
    if (j < 0 || (j & ~kMinus) > 1) { s = c.s[0]; co = c.c[0]; }
    else if constexpr (TT::kPM) c.pm(0, j & ~kMinus, j & kMinus, s, co);
    else c.sc(x[1], &s, &co);
  }
  template <class TT> __device__ static void trig2(const double* x, int j, const TT& c, double& s, double& co) {
    if (j < 0 || (j & ~kMinus) > 2) { s = c.s[1]; co = c.c[1]; }
    else if constexpr (TT::kPM) c.pm(1, j & ~kMinus, j & kMinus, s, co);
    else c.sc(x[2], &s, &co);
  }
  template <class T>
  __device__ static void dyn_g(const T* xu, const double*, T* y) {
    const double dt = 1.0 / 125.0, g = 9.81, Mc = 0.37, Mp1 = 0.127, Mp2 = 0.127, Mt = Mc + Mp1 + Mp2;
    const double L1 = 0.3365, L2 = 0.3365, l1 = L1 / 2, l2 = L2 / 2, J1 = Mp1 * L1 / 12, J2 = Mp2 * L2 / 12;
    const double a12 = Mp1 * l1 + Mp2 * L2, a13 = Mp2 * l2, a23 = L1 * l2 * Mp2;
    const double M22 = l1 * l1 * Mp1 + L1 * L1 * Mp2 + J1, M33 = l2 * l2 * Mp2 + J2;
    T s1 = t_sin(xu[1]), c1 = t_cos(xu[1]), s2 = t_sin(xu[2]), c2 = t_cos(xu[2]);
    T sd = t_sin(xu[1] - xu[2]), cd = t_cos(xu[1] - xu[2]);
    T M12 = a12 * c1, M13 = a13 * c2, M23 = a23 * cd;
    T td1 = xu[4], td2 = xu[5];
    T C12 = (-a12) * td1 * s1, C13 = (-a13) * td2 * s2, C23 = a23 * td2 * sd, C32 = (-a23) * td1 * sd;
    T G2 = (-(Mp1 * l1 + Mp2 * L1) * g) * s1, G3 = (-Mp2 * l2 * g) * s2;
    T u = 3.0 * t_clip(xu[6], -10.0, 10.0);
    T r1 = u - (C12 * td1 + C13 * td2);
    T r2 = -(C23 * td2) - G2;
    T r3 = -(C32 * td1) - G3;
    // Cholesky solve of the SPD mass matrix [[Mt, M12, M13], [M12, M22, M23], [M13, M23, M33]]
    const double l11 = sqrt(Mt);
    T l21 = M12 / l11, l31 = M13 / l11;
    T l22 = t_sqrt(M22 - l21 * l21);
    T l32 = (M23 - l31 * l21) / l22;
    T l33 = t_sqrt(M33 - l31 * l31 - l32 * l32);
    T y1 = r1 / l11;
    T y2 = (r2 - l21 * y1) / l22;
    T y3 = (r3 - l31 * y1 - l32 * y2) / l33;
    T a3 = y3 / l33;
    T a2 = (y2 - l32 * a3) / l22;
    T a1 = (y1 - l21 * a2 - l31 * a3) / l11;
    T v1 = xu[3] + a1 * dt, v2 = xu[4] + a2 * dt, v3 = xu[5] + a3 * dt;
    y[0] = xu[0] + v1 * dt;
    y[1] = xu[1] + v2 * dt;
    y[2] = xu[2] + v3 * dt;
    y[3] = v1; y[4] = v2; y[5] = v3;
  }
  __host__ __device__ static constexpr int nl_angle(int k) { return k < 2 ? 1 : 2; }
  static constexpr int OBS_NL = 4, OBS_JMAX = 2;  // z = [x, s1, c1, s2, c2, xd, thd1, thd2, u]
  __host__ __device__ static constexpr int obs_src(int a) { return a == 0 ? 0 : (a <= 4 ? -a : a - 2); }
  __host__ __device__ static constexpr int term_src(int a) { return a == 0 ? 0 : (a <= 4 ? -a : a - 2); }
  template <class TT> __device__ static void trig_nl(const double* x, int j, const TT& c, double* y) {
    trig1(x, j, c, y[0], y[1]);
    trig2(x, j, c, y[2], y[3]);
  }
  template <class TT> __device__ static void dyn(const double* xu, int j, const TT& c, const double*, double* y) {
    // kDcp = {dt, a12, a13, a23, M22, M33, -a12, -a13, -a23, -(Mp1 l1 + Mp2 L1) g, -Mp2 l2 g, 3, -10, 10, 1/sqrt(Mt)}
    const double dt = kDcp[0], a12 = kDcp[1], a13 = kDcp[2], a23 = kDcp[3], M22 = kDcp[4], M33 = kDcp[5];
    double s1, c1, s2, c2, sd, cd;
    trig1(xu, j, c, s1, c1);
    trig2(xu, j, c, s2, c2);
    if constexpr (TT::kPM) {  // sin / cos(th1 - th2) from the angles' own values: no third evaluation per point
      sd = fma(s1, c2, -(c1 * s2));
      cd = fma(c1, c2, s1 * s2);
    } else {
      c.sc(xu[1] - xu[2], &sd, &cd);
    }
    double M12 = a12 * c1, M13 = a13 * c2, M23 = a23 * cd;
    double td1 = xu[4], td2 = xu[5];
    double C12 = kDcp[6] * td1 * s1, C13 = kDcp[7] * td2 * s2, C23 = a23 * td2 * sd, C32 = kDcp[8] * td1 * sd;
    double G2 = kDcp[9] * s1, G3 = kDcp[10] * s2;
    double u = kDcp[11] * clipd(xu[6], kDcp[12], kDcp[13]);
    double r1 = u - (C12 * td1 + C13 * td2);
    double r2 = -(C23 * td2) - G2;
    double r3 = -(C32 * td1) - G3;
    // solve the SPD 3x3 system M qdd = r by Cholesky (the reference forms inv(M) @ r)
    const double i11 = kDcp[14];  // 1/sqrt(Mt), Mt = 0.624
    double l21 = M12 * i11, l31 = M13 * i11;
    double i22 = fast_rsqrt(M22 - l21 * l21);
    double l32 = (M23 - l31 * l21) * i22;
    double i33 = fast_rsqrt(M33 - l31 * l31 - l32 * l32);
    double y1 = r1 * i11;
    double y2 = (r2 - l21 * y1) * i22;
    double y3 = (r3 - l31 * y1 - l32 * y2) * i33;
    double a3 = y3 * i33;
    double a2 = (y2 - l32 * a3) * i22;
    double a1 = (y1 - l21 * a2 - l31 * a3) * i11;
    double v1 = fma(a1, dt, xu[3]), v2 = fma(a2, dt, xu[4]), v3 = fma(a3, dt, xu[5]);
    y[0] = fma(v1, dt, xu[0]);
    y[1] = fma(v2, dt, xu[1]);
    y[2] = fma(v3, dt, xu[2]);
    y[3] = v1; y[4] = v2; y[5] = v3;
  }
  template <class TT> __device__ static void obs(const double* xu, int j, const TT& c, double* z) {
    z[0] = xu[0];
    trig1(xu, j, c, z[1], z[2]);
    trig2(xu, j, c, z[3], z[4]);
    z[5] = xu[3]; z[6] = xu[4]; z[7] = xu[5]; z[8] = xu[6];
  }
  template <class TT> __device__ static void obs_term(const double* x, int j, const TT& c, double* z) {
    z[0] = x[0];
    trig1(x, j, c, z[1], z[2]);
    trig2(x, j, c, z[3], z[4]);
    z[5] = x[3]; z[6] = x[4]; z[7] = x[5];
  }
  template <class TT> __device__ static void measure(const double*, int, const TT&, double*) {}
};

// ------------------------------------------------------------------ planar quadrotor
// fp64 restatement of QuadrotorDef.step (Box2D single free body, mpc_quad.py:325-350) -- see
// oracle/envs.py:Quadrotor for the statement and DESIGN.md "parity unpinned"; measure(): mpc_quad.py:371-383
// (typos in the right-rotor velocities reproduced).  observe is the identity.
struct EnvQuadrotor {
  static constexpr int DX = 6, DU = 2, DZ = 8, DZT = 6, NP = 0, DY = 8, NA = 1;
  static constexpr bool HAS_TERM = true, LINEAR = false;
  static constexpr bool J_HANDOFF = false;  // smoother gain computed by a helper warp (i2c_kernels.cuh: JOFF)
  __host__ __device__ static constexpr double obsE(int, int) { return 0.0; }
  __host__ __device__ static constexpr double obsF(int, int) { return 0.0; }
  static constexpr int NJ = 3;
  using TrigT = Trig<NA, 0, NJ>;
  static constexpr double VDX = 0.8;                                      // W/25
  static constexpr double MASS = 5.0 * (2 * 0.8) * (2 * (400.0 / 30.0 / 100.0));
  static constexpr double INERTIA = MASS * ((2 * 0.8) * (2 * 0.8) + (2 * (400.0 / 30.0 / 100.0)) * (2 * (400.0 / 30.0 / 100.0))) / 12.0;
  template <class TT> __device__ static void center(const double* m, TT& t) { t.sc(m[2], &t.s[0], &t.c[0]); }
  template <class TT> __device__ static void offsets(const double* L, double sf, TT& t) {
    if constexpr (TT::kPM) {
#pragma unroll
      for (int j = 0; j < 3; ++j) t.sc(sf * L[tix(2, j)], &t.ds[j], &t.dc[j]);
    }
  }
  template <class TT> __device__ static void trig(const double* x, int j, const TT& c, double& s, double& co) {
    if (j < 0 || (j & ~kMinus) > 2) { s = c.s[0]; co = c.c[0]; }
    else if constexpr (TT::kPM) c.pm(0, j & ~kMinus, j & kMinus, s, co);
    else c.sc(x[2], &s, &co);
  }
  template <class T>
  __device__ static void dyn_g(const T* xu, const double*, T* y) {
    const double h = 0.1;
    T u1 = t_clip(xu[6], 0.0, 30.0), u2 = t_clip(xu[7], 0.0, 30.0);
    T s = t_sin(xu[2]), co = t_cos(xu[2]);
    T f = u1 + u2;
    T vx = xu[3] + h * ((-s * f) * (1.0 / MASS));
    T vy = xu[4] + h * (-9.81 + (co * f) * (1.0 / MASS));
    T w = xu[5] + (h * VDX / INERTIA) * (u2 - u1);
    w = w * (1.0 / (1.0 + h * 0.5));
    y[0] = xu[0] + h * vx;
    y[1] = xu[1] + h * vy;
    y[2] = xu[2] + h * w;
    y[3] = vx; y[4] = vy; y[5] = w;
  }
  __host__ __device__ static constexpr int nl_angle(int) { return 2; }
  static constexpr int OBS_NL = 0, OBS_JMAX = -1;  // observe / observe_terminal are identities
  __host__ __device__ static constexpr int obs_src(int a) { return a; }
  __host__ __device__ static constexpr int term_src(int a) { return a; }
  template <class TT> __device__ static void trig_nl(const double*, int, const TT&, double*) {}
  template <class TT> __device__ static void dyn(const double* xu, int j, const TT& c, const double*, double* y) {
    // kQuad = {h, 1/MASS, -g, VDX, 1/INERTIA, 1/(1 + h/2), 0, 30}
    const double h = kQuad[0];
    double u1 = clipd(xu[6], kQuad[6], kQuad[7]), u2 = clipd(xu[7], kQuad[6], kQuad[7]);
    double s, co;
    trig(xu, j, c, s, co);
    double f = u1 + u2;
    double vx = xu[3] + h * ((-s * f) * kQuad[1]);
    double vy = xu[4] + h * (kQuad[2] + (co * f) * kQuad[1]);
    double w = xu[5] + h * (kQuad[3] * (u2 - u1)) * kQuad[4];
    w = w * kQuad[5];
    y[0] = xu[0] + h * vx;
    y[1] = xu[1] + h * vy;
    y[2] = xu[2] + h * w;
    y[3] = vx; y[4] = vy; y[5] = w;
  }
  template <class TT> __device__ static void obs(const double* xu, int, const TT&, double* z) {
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = xu[i];
  }
  template <class TT> __device__ static void obs_term(const double* x, int, const TT&, double* z) {
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = x[i];
  }
  template <class TT> __device__ static void measure(const double* x, int j, const TT& c, double* yv) {
    double s, co;
    trig(x, j, c, s, co);
    yv[0] = x[0] - VDX * co;
    yv[1] = x[1] - VDX * s;
    yv[2] = x[0] + VDX * co;
    yv[3] = x[1] + VDX * s;
    yv[4] = x[3] - VDX * -s * x[5];
    yv[5] = x[4] - VDX * co * x[5];
    yv[6] = x[3] + VDX - s * x[5];
    yv[7] = x[4] + VDX + co * x[5];
  }
};

}  // namespace i2c
