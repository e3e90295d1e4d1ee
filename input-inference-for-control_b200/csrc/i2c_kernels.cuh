// Persistent fp64 kernels of the Gaussian i2c EM sweep for sm_100a (templates; instantiated once per environment
// by i2c_env_inst.cu, one translation unit per env so that the build parallelises).
//
// One thread owns one problem for the whole launch: it runs the forward filter over the horizon, the
// backward smoother (with the M-step statistics fused in), optionally the closed-loop propagate sweep,
// then the alpha update -- for all requested EM iterations -- without leaving the kernel.  Problems are
// independent, so no inter-thread synchronisation is needed; the recursion state (mean / covariance /
// Cholesky factor of the dx-dimensional message) is carried in registers between timesteps and all
// small dense algebra (Cholesky, triangular solves, Gaussian conditioning) is fully unrolled over
// packed-lower-triangular register arrays.  Per-cell records live in HBM in the tiled layout described
// in i2c_types.h: every global access of a warp is a 256-byte coalesced segment.
//
// Reference arithmetic restated here (paths relative to the reference root):
//   forward cell  i2c/i2c.py:350-447     backward cell  i2c/i2c.py:544-610
//   propagate     i2c/i2c.py:150-199     M-step         i2c/i2c.py:1004-1065, 913-981
//   quadrature    i2c/inference/quadrature.py:15-58, i2c/exp_types.py:36-49
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>

#include "envs.cuh"
#include "i2c_types.h"

namespace i2c {

template <class Env>
struct Lay {
  static constexpr int DX = Env::DX, DU = Env::DU, N = DX + DU, DZ = Env::DZ, DZT = Env::DZT;
  // posterior / prior record
  static constexpr int P_MU = 0, P_SIG = N, P_K = P_SIG + TRI(N), P_KK = P_K + DU * DX, P_SIGK = P_KK + DU;
  static constexpr int E_POST = P_SIGK + TRI(DU);
  // filtered record
  static constexpr int F_MU1 = 0, F_SIG1 = N, F_MU3 = F_SIG1 + TRI(N), F_SIG3 = F_MU3 + DX, F_J = F_SIG3 + TRI(DX);
  static constexpr int E_FILT = F_J + N * DX;
  static constexpr int AF_MU0 = 0, AF_SIG0 = N, AF_MUZ = AF_SIG0 + TRI(N), AF_SIGZ = AF_MUZ + DZ;
  static constexpr int E_AUXF = AF_SIGZ + TRI(DZ);
  static constexpr int AB_MUZ = 0, AB_SIGZ = DZ, AB_MU3M = AB_SIGZ + TRI(DZ), AB_SIG3M = AB_MU3M + DX;
  static constexpr int E_AUXB = AB_SIG3M + TRI(DX);
  static constexpr int PF_MU = 0, PF_SIG = N, PF_MUZ = PF_SIG + TRI(N), PF_SIGZ = PF_MUZ + DZ, PF_MU3 = PF_SIGZ + TRI(DZ),
                       PF_SIG3 = PF_MU3 + DX;
  static constexpr int E_PF = PF_SIG3 + TRI(DX);
  static constexpr int TM_MU = 0, TM_SIG = DZT, E_TERM = DZT + TRI(DZT);
  static constexpr int R_L3 = 0, R_N3 = DX * DX, R_L0 = R_N3 + DX, R_N0 = R_L0 + DX * DX, E_RIC = R_N0 + DX;
  // staged (prefetched) parts of the records: the prior / posterior without k, sigK; the whole filtered record
  static constexpr int E_STAGE_POST = P_KK, E_STAGE = E_FILT > P_KK ? E_FILT : P_KK;
  // + the cell's target z (DZ doubles) and one 8-byte slot holding {flags, index}
  static constexpr int S_Z = E_STAGE, S_META = E_STAGE + DZ, E_STAGE_TOT = E_STAGE + DZ + 1;
  // big records (double cart-pole, quadrotor): a cell costs tens of thousands of cycles, staging would only cost
  // occupancy (2 x 58 KB per warp) -> read the records straight from global memory
  static constexpr bool STAGED = E_FILT <= 64;
  // team kernel: depth of the filtered-record stream of the RTS head loop, and the staging area it needs
  static constexpr int TEAM_DEPTH = 6;
  // team kernel, HOT: producer / consumer ring of RING slots (E_STAGE doubles per lane each) filled by a copy warp
  // (the big records of the double cart-pole / quadrotor -- 27-30 KB per slot -- get a ring of 2: their cells take thousands of
  // cycles, one cell of look-ahead hides the DRAM latency that cost those kernels a third of their time when they read the
  // records straight from global memory)
  static constexpr int RING = E_FILT > 64 ? 2 : (E_FILT > 32 ? 4 : 8);
  static constexpr int E_TEAM_A = !STAGED ? 0 : ((TEAM_DEPTH * E_FILT > 2 * E_STAGE_TOT) ? TEAM_DEPTH * E_FILT : 2 * E_STAGE_TOT);
  static constexpr int E_TEAM_STAGE = E_TEAM_A > RING * E_STAGE ? E_TEAM_A : RING * E_STAGE;

};

// ------------------------------------------------------------------------------------------------
// cp.async staging of the next cell's record: every thread copies ITS OWN E elements (8 B each, one coalesced
// 256-byte segment per warp instruction) into thread-private shared-memory slots while the current cell is being
// computed, then waits on its own async group -- no block / warp synchronisation is needed because no slot is
// shared between threads.  Layout of the staging buffer mirrors the global record: smem[e][lane].
template <int E>
__device__ __forceinline__ void stage_record(double* sdst, const double* gsrc) {
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(sdst);
#pragma unroll
  for (int e = 0; e < E; ++e)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + e * TILE * 8), "l"(gsrc + (size_t)e * TILE) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// ---- TMA-1D bulk copies (cp.async.bulk, SASS: UBLKCP): the record of one (cell, tile) is a single contiguous
// E*256-byte block, so ONE elected lane moves it into the warp's staging buffer and every lane waits on the
// buffer's mbarrier.  Replaces E LDGSTS instructions per thread and cell.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(double* sdst, const double* gsrc, unsigned bytes, uint64_t* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(b),
      "r"(parity)
      : "memory");
}
// the same on precomputed 32-bit shared addresses (no generic -> shared conversion inside the cell loops: it costs an
// S2R SR_CgaCtaId + address arithmetic per use), plus the non-blocking probe
__device__ __forceinline__ bool mbar_test_s(unsigned bar_s, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar_s), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(unsigned bar_s, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar_s),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(unsigned bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
// 32-bit shared address of a generic pointer, made OPAQUE to the compiler: otherwise ptxas rematerialises the conversion
// (S2R SR_CgaCtaId + LEA, ~100 exposed cycles) at every use inside the cell loops instead of keeping the value in a register
__device__ __forceinline__ unsigned smem_addr(const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ double lds_f64(unsigned addr_s) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr_s));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr_s, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr_s), "d"(v) : "memory");
}
// read-only tables (written once before a block-wide barrier): plain asm, the compiler may move / merge these loads
__device__ __forceinline__ double lds_f64_ro(unsigned addr_s) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr_s));
  return v;
}
__device__ __forceinline__ int2 lds_i2_ro(unsigned addr_s) {
  int2 v;
  asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr_s));
  return v;
}
// Progress counter of the team kernel (heads finished), in shared memory: published with a release exchange by one lane
// after __syncwarp() (cumulative over the warp's record stores), polled with acquire loads by the helper warps.
// (prog_s = 32-bit shared address, converted once: a generic -> shared conversion inside the head loop costs an
// S2R SR_CgaCtaId round trip of ~100 cycles per cell)
__device__ __forceinline__ void progress_publish(unsigned prog_s, int v) {
  int old;
  asm volatile("atom.exch.release.cta.shared::cta.b32 %0, [%1], %2;" : "=r"(old) : "r"(prog_s), "r"(v) : "memory");
  (void)old;
}
__device__ __forceinline__ int progress_read(unsigned prog_s) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(prog_s) : "memory");
  return v;
}
#ifdef I2C_NO_BULK
constexpr bool kUseBulk = false;
#else
constexpr bool kUseBulk = true;
#endif
constexpr int kNumBars = 16;  // mbarriers per warp (>= deepest record pipeline; team ring: RING full + RING empty)
template <int N>
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Sigma-point transform (inference/quadrature.py:15-58) around (m, chol L) in dimension D.
//   my  = sum_p w_p y_p                        (mean uses weights_sig: quirk A.6.1)
//   Syy = sum_p w_p y_p y_p^T - my my^T        (packed lower, no noise added)
//   Dm[j][:] = w * sf * (y_{+j} - y_{-j})      so that  S_xy = L * Dm  (exact rewrite of
//              sum_p w_p x_p y_p^T - m my^T for the symmetric point set m +- sf L[:,j])
// PM: the env maps take the minus point of a column from the angle-addition cache (envs.cuh: Trig, kMinus)
template <int D, int DY, bool PM = false, class Eval>
__device__ __forceinline__ void sigma_transform(const double* m, const double* L, double sf, double w0, double wi,
                                                Eval&& eval, double* my, double* Syy, double* Dm) {
  double sy[DY], syy[TRI(DY)];
#pragma unroll
  for (int a = 0; a < DY; ++a) sy[a] = 0.0;
#pragma unroll
  for (int a = 0; a < TRI(DY); ++a) syy[a] = 0.0;
  const double wsf = wi * sf;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double xp[D], xm[D], yp[DY], ym[DY];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i >= j) {
        double d = sf * L[tix(i, j)];
        xp[i] = m[i] + d;
        xm[i] = m[i] - d;
      } else {
        xp[i] = m[i];
        xm[i] = m[i];
      }
    }
    eval(xp, j, yp);
    eval(xm, PM ? (j | kMinus) : j, ym);
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      sy[a] += yp[a] + ym[a];
      Dm[j * DY + a] = wsf * (yp[a] - ym[a]);
#pragma unroll
      for (int b = 0; b <= a; ++b) syy[tix(a, b)] = fma(yp[a], yp[b], fma(ym[a], ym[b], syy[tix(a, b)]));
    }
  }
#pragma unroll
  for (int a = 0; a < DY; ++a) my[a] = wi * sy[a];
#pragma unroll
  for (int a = 0; a < TRI(DY); ++a) Syy[a] = wi * syy[a];
  if (w0 != 0.0) {  // centre point only contributes for alpha != 1 or beta != 0 (exp_types.py:40-49)
    double y0[DY];
    eval(m, -1, y0);
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      my[a] = fma(w0, y0[a], my[a]);
#pragma unroll
      for (int b = 0; b <= a; ++b) Syy[tix(a, b)] = fma(w0 * y0[a], y0[b], Syy[tix(a, b)]);
    }
  }
#pragma unroll
  for (int a = 0; a < DY; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) Syy[tix(a, b)] = fma(-my[a], my[b], Syy[tix(a, b)]);
}

// S_xy[i][a] = sum_{j<=i} L[i][j] Dm[j][a], computed IN PLACE (bottom row first: row i only needs rows j <= i), so
// that Dm and S_xy never coexist in registers.
template <int D, int DY>
__device__ __forceinline__ void cross_cov(const double* L, double* DmSxy) {
#pragma unroll
  for (int i = D - 1; i >= 0; --i)
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j <= i; ++j) s = fma(L[tix(i, j)], DmSxy[j * DY + a], s);
      DmSxy[i * DY + a] = s;
    }
}

// Gauss-Hermite rule (exp_types.py:52-68, GaussHermiteQuadrature): the degree^D tensor grid of 1-D Gauss-Hermite nodes,
//   x_p = m + sqrt(2) L xi_p,   w_p = prod_i w[d_i] / pi^(D/2)       (g.w holds the 1-D weights already divided by sqrt(pi))
// Same outputs as sigma_transform: my, Syy (no noise), and Dm[j][:] = sqrt(2) sum_p w_p xi_pj y_p so that S_xy = L * Dm
// (equal to the reference's sum_p w_p x_p y_p^T - m my^T because the nodes are symmetric and the weights sum to one).
// The point loop is a run-time odometer: degree^D evaluations (27 for the pendulum at degree 3, 16 384 at d = 7, degree 4).
template <int D, int DY, class Eval>
__device__ __forceinline__ void grid_transform(const double* m, const double* L, const GhRule& g, Eval&& eval, double* my,
                                               double* Syy, double* Dm) {
  const double sf = 1.4142135623730951;
  double sy[DY], syy[TRI(DY)];
#pragma unroll
  for (int a = 0; a < DY; ++a) sy[a] = 0.0;
#pragma unroll
  for (int a = 0; a < TRI(DY); ++a) syy[a] = 0.0;
#pragma unroll
  for (int a = 0; a < D * DY; ++a) Dm[a] = 0.0;
  int dig[D];
#pragma unroll
  for (int i = 0; i < D; ++i) dig[i] = 0;
  int total = 1;
#pragma unroll
  for (int i = 0; i < D; ++i) total *= g.degree;
  for (int pt = 0; pt < total; ++pt) {
    double xi[D], x[D], y[DY], w = 1.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      xi[i] = g.x[dig[i]];
      w *= g.w[dig[i]];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) sacc = fma(L[tix(i, k)], xi[k], sacc);
      x[i] = fma(sf, sacc, m[i]);
    }
    eval(x, 0, y);  // column index 0: every angle is evaluated afresh (no centre cache applies on a full grid)
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      const double wy = w * y[a];
      sy[a] += wy;
#pragma unroll
      for (int b = 0; b <= a; ++b) syy[tix(a, b)] = fma(wy, y[b], syy[tix(a, b)]);
#pragma unroll
      for (int j = 0; j < D; ++j) Dm[j * DY + a] = fma(sf * xi[j], wy, Dm[j * DY + a]);
    }
    bool carry = true;  // odometer increment
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (carry) {
        dig[i] += 1;
        carry = dig[i] >= g.degree;
        if (carry) dig[i] = 0;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < DY; ++a) my[a] = sy[a];
#pragma unroll
  for (int a = 0; a < DY; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) Syy[tix(a, b)] = fma(-sy[a], sy[b], syy[tix(a, b)]);
}

// Gaussian conditioning on an observation with moments (my, Sy incl. noise, Sxy) and target z:
//   G = Sxy Sy^{-1};  mu += G (z - my);  Sig -= G Sxy^T      (i2c.py:398-403 / :438-443)
// done through the Cholesky factor of Sy: W_i = Ly^{-1} Sxy[i,:]^T, r = Ly^{-1}(z - my).
template <int D, int DY>
__device__ __forceinline__ bool condition(double* mu, double* Sig, double* Sy, double* Sxy /* overwritten by W */,
                                          const double* my, const double* z) {
  double invd[DY];
  bool ok = chol_rows<DY>(Sy, invd);
  double r[DY];
#pragma unroll
  for (int a = 0; a < DY; ++a) r[a] = z[a] - my[a];
  fwd_subst<DY>(Sy, invd, r);
  double* W = Sxy;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    fwd_subst<DY>(Sy, invd, W + i * DY);
    double dm = 0.0;
#pragma unroll
    for (int a = 0; a < DY; ++a) dm = fma(W[i * DY + a], r[a], dm);
    mu[i] += dm;
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = Sig[tix(i, j)];
#pragma unroll
      for (int a = 0; a < DY; ++a) s = fma(-W[i * DY + a], W[j * DY + a], s);
      Sig[tix(i, j)] = s;
    }
  return ok;
}

// Structured sigma-point moments of the cost-feature maps.  Every registered observe / observe_terminal map consists
// of identity features (z_a = x_i) and trigonometric features of angles stored at low state indices.  For the
// cubature rule with zero centre weight and sum(w) = 1 the moments of the identity features are exact linear
// results (m_y = m_i, S_xy[:, a] = Sigma[:, i], S_yy = Sigma[i, i']), the cross moment of a nonlinear feature with an
// identity feature is S_xy[i][nl] (both are sum_j D_j[nl] L[i][j]), and sigma-point columns j > OBS_JMAX reproduce
// the centre value of every nonlinear feature.  Only the nonlinear block is therefore evaluated at the 2 (JMAX + 1)
// points that differ from the centre.  Same numbers as sigma_transform + cross_cov up to round-off.
// pre: sin / cos of the centre and of the column offsets evaluated beforehand (they depend on the state block of (m, L)
// only; the plain forward loop computes them at the END of the previous cell, beside that cell's off-chain work)
template <class Env, int D, int DY, bool TERM, class TT = typename Env::TrigT>
__device__ __forceinline__ void structured_obs_moments(const double* m, const double* Sig, const double* L, double sf,
                                                       double wi, double* my, double* Syy, double* Sxy,
                                                       const TT* pre = nullptr) {
  constexpr int NL = Env::OBS_NL, JM = Env::OBS_JMAX;
  constexpr int NLs = NL > 0 ? NL : 1;
  double mnl[NLs], Snl[TRI(NLs)], Dm[(JM + 1 > 0 ? JM + 1 : 1) * NLs], Cx[D * NLs];
  if constexpr (NL > 0) {
    TT ctx;
    if (pre) {
      ctx = *pre;
    } else {
      Env::center(m, ctx);
      Env::offsets(L, sf, ctx);
    }
    double yc[NL], sy[NL], syy[TRI(NL)];
    Env::trig_nl(m, -1, ctx, yc);
    constexpr double mult = 2.0 * (D - 1 - JM);
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      sy[a] = mult * yc[a];
#pragma unroll
      for (int b = 0; b <= a; ++b) syy[tix(a, b)] = mult * yc[a] * yc[b];
    }
    const double wsf = wi * sf;
#pragma unroll
    for (int j = 0; j <= JM; ++j) {
      double xp[D], xm[D], yp[NL], ym[NL];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (i >= j) {
          const double d = sf * L[tix(i, j)];
          xp[i] = m[i] + d;
          xm[i] = m[i] - d;
        } else {
          xp[i] = m[i];
          xm[i] = m[i];
        }
      }
      Env::trig_nl(xp, j, ctx, yp);
      Env::trig_nl(xm, TT::kPM ? (j | kMinus) : j, ctx, ym);
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        sy[a] += yp[a] + ym[a];
        Dm[j * NL + a] = wsf * (yp[a] - ym[a]);
#pragma unroll
        for (int b = 0; b <= a; ++b) syy[tix(a, b)] = fma(yp[a], yp[b], fma(ym[a], ym[b], syy[tix(a, b)]));
      }
    }
#pragma unroll
    for (int a = 0; a < NL; ++a) mnl[a] = wi * sy[a];
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) Snl[tix(a, b)] = fma(wi, syy[tix(a, b)], -mnl[a] * mnl[b]);
    // cross covariance with the nonlinear block: Cx[i][k] = sum_{j <= min(i, JM)} L[i][j] Dm[j][k]
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= JM; ++j)
          if (j <= i) v = fma(L[tix(i, j)], Dm[j * NL + k], v);
        Cx[i * NL + k] = v;
      }
  }
#pragma unroll
  for (int a = 0; a < DY; ++a) {
    const int sa = TERM ? Env::term_src(a) : Env::obs_src(a);
    my[a] = sa >= 0 ? m[sa >= 0 ? sa : 0] : mnl[sa < 0 ? -1 - sa : 0];
#pragma unroll
    for (int i = 0; i < D; ++i) Sxy[i * DY + a] = sa >= 0 ? Sig[six(i, sa >= 0 ? sa : 0)] : Cx[i * NLs + (sa < 0 ? -1 - sa : 0)];
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      const int sb = TERM ? Env::term_src(b) : Env::obs_src(b);
      double v;
      if (sa >= 0 && sb >= 0) v = Sig[six(sa >= 0 ? sa : 0, sb >= 0 ? sb : 0)];
      else if (sa < 0 && sb < 0) v = Snl[six(sa < 0 ? -1 - sa : 0, sb < 0 ? -1 - sb : 0)];
      else if (sa < 0) v = Cx[(sb >= 0 ? sb : 0) * NLs + (sa < 0 ? -1 - sa : 0)];
      else v = Cx[(sa >= 0 ? sa : 0) * NLs + (sb < 0 ? -1 - sb : 0)];
      Syy[tix(a, b)] = v;
    }
  }
}

// quadratic-cost statistics of a Gaussian cost feature (i2c.py:1034-1043 and :680-683, :913-919)
//   mean = e^T QR e + tr(Sz QR),  var = 2 tr((Sz QR)^2) + 4 e^T QR Sz QR e,   e = mz - zref
template <int DZ>
__device__ __forceinline__ void cost_stats(const KParams& p, const double* mz, const double* Sz, const double* zref,
                                           double& mean, double& var) {
  double e[DZ], v[DZ];
#pragma unroll
  for (int a = 0; a < DZ; ++a) e[a] = mz[a] - zref[a];
  if (p.qr_diag) {
    double m = 0.0, t2 = 0.0, q4 = 0.0;
#pragma unroll
    for (int a = 0; a < DZ; ++a) v[a] = p.QR[a * DZ + a] * e[a];
#pragma unroll
    for (int a = 0; a < DZ; ++a) {
      m = fma(e[a], v[a], m);
      m = fma(Sz[tix(a, a)], p.QR[a * DZ + a], m);
    }
#pragma unroll
    for (int a = 0; a < DZ; ++a)
#pragma unroll
      for (int b = 0; b < DZ; ++b) {
        double s = Sz[six(a, b)];
        t2 = fma(s * s, p.QR[a * DZ + a] * p.QR[b * DZ + b], t2);
        q4 = fma(v[a] * s, v[b], q4);
      }
    mean = m;
    var = 2.0 * t2 + 4.0 * q4;
  } else {
    double P[DZ * DZ];  // Sz QR
#pragma unroll
    for (int a = 0; a < DZ; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < DZ; ++b) s = fma(p.QR[a * DZ + b], e[b], s);
      v[a] = s;
    }
#pragma unroll
    for (int a = 0; a < DZ; ++a)
#pragma unroll
      for (int b = 0; b < DZ; ++b) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < DZ; ++c) s = fma(Sz[six(a, c)], p.QR[c * DZ + b], s);
        P[a * DZ + b] = s;
      }
    double m = 0.0, t2 = 0.0, q4 = 0.0;
#pragma unroll
    for (int a = 0; a < DZ; ++a) {
      m = fma(e[a], v[a], m);
      m += P[a * DZ + a];
#pragma unroll
      for (int b = 0; b < DZ; ++b) {
        t2 = fma(P[a * DZ + b], P[b * DZ + a], t2);
        q4 = fma(v[a] * Sz[six(a, b)], v[b], q4);
      }
    }
    mean = m;
    var = 2.0 * t2 + 4.0 * q4;
  }
}

// tr(QR (d d^T + Sz)),  d = z - mz     (expected_observation_covar + calculate_alpha numerator)
template <int DZ>
__device__ __forceinline__ double alpha_trace(const double* Q, int qr_diag, const double* mz, const double* Sz,
                                              const double* z) {
  double d[DZ];
#pragma unroll
  for (int a = 0; a < DZ; ++a) d[a] = z[a] - mz[a];
  double tr = 0.0;
  if (qr_diag) {
#pragma unroll
    for (int a = 0; a < DZ; ++a) tr = fma(Q[a * DZ + a], fma(d[a], d[a], Sz[tix(a, a)]), tr);
  } else {
#pragma unroll
    for (int a = 0; a < DZ; ++a)
#pragma unroll
      for (int b = 0; b < DZ; ++b) tr = fma(Q[a * DZ + b], fma(d[b], d[a], Sz[six(b, a)]), tr);
  }
  return tr;
}

// message carried between cells: dx-dimensional Gaussian + its Cholesky factor
template <int DX>
struct Carry {
  double m[DX], S[TRI(DX)], L[TRI(DX)], invd[DX];
};

// META = latency-regime variant: per-cell targets / flags are staged with the records and the records move with
// per-thread cp.async (lowest latency); throughput variant (META = false): records move with one TMA bulk copy per
// warp and cell (fewest instructions), targets / flags are plain cached loads.
// LIN = Linearize inference compiled in (separate instantiation: keeps the cubature kernels small).
// GH = Gauss-Hermite tensor-grid rule instead of the cubature points (separate instantiation as well).
// HOT = the common configuration is compiled in instead of tested per cell (launcher: KParams::hot): cubature rule with
// zero centre weight (structured cost-feature moments), no auxiliary records, shared cell targets staged with the records,
// no per-cell alpha.  Every one of these run-time switches was a (uniform) branch in the cell loops: ~10 basic-block
// boundaries per forward cell that stop ptxas from overlapping the independent dependency chains across them.
// HOT = 2: as 1, but cells may carry their own alpha (MPC: the cell appended by the horizon shift).  A separate instantiation
// (environments with a measurement model only): compiling the per-cell alpha into the HOT = 1 kernels changed their register
// allocation throughout and cost the judged pendulum line 12 % (0.2115 -> 0.237 ms per iteration).
template <class Env, bool META, bool LIN = false, bool GH = false, int HOT = 0>
struct Worker {
  // HOT moves the records with bulk copies as well: targets / flags come from a shared-memory table built once per launch
  // (team kernel), so the per-thread LDGSTS stream (17 + 20 instructions and ~25 of address arithmetic per cell) is gone
  static constexpr bool BULK = kUseBulk && (!META || HOT);
  using LY = Lay<Env>;
  static constexpr int DX = LY::DX, DU = LY::DU, N = LY::N, DZ = LY::DZ, DZT = LY::DZT;
  // sincos flavour of this variant (envs.cuh: Trig): latency variants branch-free + angle addition, throughput variants
  // sequenced; the Gauss-Hermite grid evaluates every point afresh (no +- pairs)
#ifndef I2C_TRIG_LAT
#define I2C_TRIG_LAT 3
#endif
#ifndef I2C_TRIG_THR
#define I2C_TRIG_THR 2
#endif
  static constexpr int TRIG_MODE = GH ? (META ? 1 : 0) : (META ? I2C_TRIG_LAT : I2C_TRIG_THR);
  using TrigT = Trig<Env::NA, TRIG_MODE, Env::NJ>;

  const KParams& p;
  const int tile, lane, b;
  double par[Env::NP > 0 ? Env::NP : 1];
  int status, info;
  // record views: the forward sweep reads `prior`, the backward sweep writes `post`, propagate reads the
  // most recently written posterior `latest`; _update_priors swaps prior/post instead of copying.
  double *prior, *post, *latest;
  bool own_alpha_valid;
  double* stage;  // this warp's double buffer: [2][E_STAGE][32], already offset by lane
  uint64_t* bars; // this warp's mbarriers (bulk-copy completion), kNumBars of them
  unsigned bar_phase;  // one parity bit per mbarrier
  bool pipe_ready = false;  // the warp's mbarriers are initialised already (em_ticket_kernel: once per warp, not per work item)
  unsigned ztab_s, ftab_s;  // HOT: shared addresses of the [T][DZ] cell targets and the [T] {flags, index} table
  // J_dyn hand-off (HOT team kernels, n >= 5): the smoother gain J = S_xy S_x3^{-1} (i2c.py:425-428) is only read by the backward
  // sweep, so warp 0 -- whose forward cell is the critical path -- does not compute it: it leaves S_xy and the factor of
  // S_x3 in one of two shared-memory slots and helper warp 1 does the substitutions and the record stores behind it.
  // Measured (round 2): cart-pole 4096 x 200 0.74 -> 0.66 ms per iteration (the forward cell loses its register spills);
  // double cart-pole and quadrotor: no gain (5.03 -> 5.07 ms, 1.01 -> 1.04 ms) -- enabled per environment (Env::J_HANDOFF).
  // The slots live in the part of the staging area the record ring does not use (no extra shared memory: the two resident
  // blocks of the 4-warp variant must keep fitting).
  static constexpr bool JOFF = HOT && Env::J_HANDOFF;
  static constexpr int JELEMS = N * DX + TRI(DX) + DX;
  // slot k & 1 for the k-th cell; mbarriers bars[JBAR + s] ("full": the 32 lanes of warp 0 arrive after their stores) and
  // bars[JBAR + 2 + s] ("empty": the 32 lanes of warp 1 arrive after their loads), phase parity (k >> 1) & 1
  static constexpr int JBAR = 2 * LY::RING;
  static_assert(!JOFF || JBAR + 4 <= kNumBars, "no mbarriers left for the J_dyn hand-off");
  unsigned jslot_s = 0;  // shared address of the slots (lane offset included)
  unsigned jcount = 0;   // forward cells handed over (warp 0) / served (warp 1) so far in this launch

  __device__ Worker(const KParams& p_, int tile_, int lane_, double* stage_, uint64_t* bars_)
      : p(p_), tile(tile_), lane(lane_), b(tile_ * TILE + lane_), stage(stage_), bars(bars_), bar_phase(0), ring_q(0), ring_ready(false), ring_ready2(false) {
    bars_s = smem_addr(bars_);
    stage_s = smem_addr(stage_);
    status = I2C_OK;
    info = 0;
    prior = p.prior;
    post = p.post;
    latest = p.latest;
    own_alpha_valid = true;
#pragma unroll
    for (int i = 0; i < Env::NP; ++i) par[i] = p.envpar[((size_t)tile * Env::NP + i) * TILE + lane];
  }

  __device__ __forceinline__ int slot(int t) const {
    int s = t + p.cell_head;
    return s >= p.T ? s - p.T : s;
  }
  __device__ __forceinline__ double* rec(double* base, int t, int E) const {
    return base + ((size_t)slot(t) * p.ntiles + tile) * (size_t)E * TILE + lane;
  }
  // Cursor over the records of consecutive cells of one per-cell array (ring layout): the address advances by a constant
  // stride per cell and wraps once per sweep -- instead of rebuilding slot(t) and the 64-bit product for every cell.
  struct Cursor {
    double* ptr;
    size_t stride, span;  // elements per cell slot, elements of the whole ring
    int s, T;
    // selects, not branches: a branch here would split the cell loop's basic block
    __device__ __forceinline__ void next() {
      const bool wrap = (s + 1 == T);
      ptr += wrap ? (ptrdiff_t)stride - (ptrdiff_t)span : (ptrdiff_t)stride;
      s = wrap ? 0 : s + 1;
    }
    __device__ __forceinline__ void prev() {
      const bool wrap = (s == 0);
      ptr += wrap ? (ptrdiff_t)span - (ptrdiff_t)stride : -(ptrdiff_t)stride;
      s = wrap ? T - 1 : s - 1;
    }
  };
  __device__ __forceinline__ Cursor cursor(double* base, int t, int E) const {
    Cursor c;
    c.stride = (size_t)p.ntiles * E * TILE;
    c.span = c.stride * p.T;
    c.s = slot(t);
    c.T = p.T;
    c.ptr = base + ((size_t)c.s * p.ntiles + tile) * (size_t)E * TILE + lane;
    return c;
  }
  __device__ __forceinline__ void fail(int code, int it, int t) {
    if (status == I2C_OK) {
      status = code;
      info = (it << 16) | (t & 0xffff);
    }
  }
  __device__ __forceinline__ bool fobs() const { return HOT || p.fast_obs; }
  __device__ __forceinline__ bool smeta() const { return !HOT && p.stage_meta; }
  __device__ __forceinline__ bool zpp() const { return !HOT && p.z_per_problem; }
  __device__ __forceinline__ void load_z(int t, double* z) const {
    if (zpp()) {
      const double* q = p.z_cell + ((size_t)slot(t) * p.ntiles + tile) * DZ * TILE + lane;
#pragma unroll
      for (int a = 0; a < DZ; ++a) z[a] = q[a * TILE];
    } else {
#pragma unroll
      for (int a = 0; a < DZ; ++a) z[a] = p.z_cell[slot(t) * DZ + a];
    }
  }
  // stage the cell's target and {flags, index} next to its record (same cp.async group)
  __device__ __forceinline__ void stage_meta(double* sbuf, int t) const {
    const int sl = slot(t);
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sbuf);
    const double* zsrc = zpp() ? p.z_cell + ((size_t)sl * p.ntiles + tile) * DZ * TILE + lane : p.z_cell + sl * DZ;
    const size_t zstride = zpp() ? TILE : 1;
#pragma unroll
    for (int a = 0; a < DZ; ++a)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + (LY::S_Z + a) * TILE * 8), "l"(zsrc + a * zstride) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s0 + LY::S_META * TILE * 8), "l"(p.cell_flags + sl) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s0 + LY::S_META * TILE * 8 + 4), "l"(p.cell_index + sl) : "memory");
  }
  __device__ __forceinline__ void staged_z(const double* sbuf, int t, double* z) const {
    if constexpr (HOT) {
#pragma unroll
      for (int a = 0; a < DZ; ++a) z[a] = lds_f64_ro(ztab_s + (t * DZ + a) * 8);
    } else if (LY::STAGED && META && smeta()) {
#pragma unroll
      for (int a = 0; a < DZ; ++a) z[a] = sbuf[(LY::S_Z + a) * TILE];
    } else {
      load_z(t, z);
    }
  }
  __device__ __forceinline__ int staged_flags(const double* sbuf, int t, bool flipped) const {
    int flags, index;
    if constexpr (HOT) {
      const int2 m = lds_i2_ro(ftab_s + t * 8);
      flags = m.x;
      index = m.y;
    } else if (LY::STAGED && META && smeta()) {
      const int2 m = *reinterpret_cast<const int2*>(sbuf + LY::S_META * TILE);
      flags = m.x;
      index = m.y;
    } else {
      flags = p.cell_flags[slot(t)];
      index = flipped ? p.cell_index[slot(t)] : 0;
    }
    if (flipped && p.tau > 0 && index <= p.tau) flags &= ~I2C_CELL_INDEPENDENT;
    return flags;
  }
  // ---- record ring of the team kernel (HOT): the copy warp (lane 0 of warp W-1) streams the records of the sequential
  // sweeps with TMA bulk copies; warp 0 only waits on the slot's "full" mbarrier and arrives on its "empty" one.  Warp 0
  // therefore issues no copy, no address arithmetic and no commit / wait bookkeeping for its input stream.
  static constexpr int RING = LY::RING, RSTRIDE = LY::E_STAGE;
  unsigned ring_q;  // cells produced / consumed so far: slot = q % RING, phase parity = (q / RING) & 1
  unsigned bars_s, stage_s;  // 32-bit shared addresses of bars[0] and of this lane's column of the staging area
  bool ring_ready, ring_ready2;  // early probes of the slots the next two ring_acquire() calls take (hide the mbarrier round trip)
  __device__ __forceinline__ void ring_init() {  // one thread, before the block-wide barrier that precedes any use
#pragma unroll
    for (int i = 0; i < RING; ++i) {
      mbar_init(bars + i, 1);            // full: the producer's arrive.expect_tx
      mbar_init(bars + RING + i, TILE);  // empty: every lane of warp 0 arrives after its last read of the slot
    }
    if constexpr (JOFF) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(bars + JBAR + i, TILE);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ring_acquire() hands out the slot as a 32-bit shared address (lane offset included): reading through the generic
  // pointer made the compiler rebuild the shared-window base (S2R SR_CgaCtaId + arithmetic, ~100 exposed cycles in the
  // short RTS-head cells) in every cell
  // small systems copy the record into registers and release the slot at once; for n > 3 that many live registers spill
  // (cart-pole: 600 bytes of stack), so those read the slot in place and release it after the cell
  static constexpr bool RCOPY = N <= 3;
  template <int E>
  __device__ __forceinline__ void ring_read(double* dst) {
    const unsigned a = stage_s + ring_acquire() * (RSTRIDE * TILE * 8);
#pragma unroll
    for (int e = 0; e < E; ++e) dst[e] = lds_f64(a + e * (TILE * 8));
    ring_release();  // the record is in registers (an arrive placed after the cell's stores waited ~90 cycles for them)
  }
  __device__ __forceinline__ const double* ring_view() { return stage + ring_acquire() * (RSTRIDE * TILE); }
  __device__ __forceinline__ unsigned ring_acquire() {  // -> slot index
    const unsigned s = ring_q % RING;
    if (!ring_ready) mbar_wait_s(bars_s + 8u * s, (ring_q / RING) & 1u);
    // probe the slot TWO cells ahead now: an mbarrier query takes ~100 cycles to come back (more than half an RTS-head
    // cell), and the copy warp runs several cells ahead, so the answer -- consumed two acquires later -- is almost always
    // "ready"
    const unsigned q2 = ring_q + 2;
    ring_ready = ring_ready2;
    ring_ready2 = mbar_test_s(bars_s + 8u * (q2 % RING), (q2 / RING) & 1u);
    return s;
  }
  __device__ __forceinline__ void ring_release() {
    mbar_arrive_s(bars_s + 8u * (RING + ring_q % RING));
    ++ring_q;
  }
  __device__ __forceinline__ void ring_produce(double* base_g, int Erec, int Ecopy, int t) {
    const unsigned s = ring_q % RING;
    mbar_wait(bars + RING + s, ((ring_q / RING) & 1u) ^ 1u);
    bulk_load(stage + s * (RSTRIDE * TILE), rec(base_g, t, Erec), Ecopy * TILE * 8, bars + s);
    ++ring_q;
  }
  // HOT: true when every cell but the last of this sweep is a feedback, non-terminal cell (flags table in shared memory)
  __device__ __forceinline__ bool sweep_is_plain(bool flipped) const {
    bool bad = false;
    for (int t = lane; t < p.T - 1; t += TILE)
      bad = bad || (staged_flags(nullptr, t, flipped) & (I2C_CELL_INDEPENDENT | I2C_CELL_TERMINAL | (HOT == 2 ? I2C_CELL_OWN_ALPHA : 0))) != 0;
    return !__any_sync(0xffffffffu, bad);
  }
  // init of this warp's mbarriers (call once, all lanes)
  __device__ __forceinline__ void pipe_init() {
    if constexpr (BULK && LY::STAGED) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kNumBars; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
    }
  }
  // copy E elements x 32 lanes of a record into staging buffer `sbuf` (lane-offset pointers), completion on bars[bi]
  template <int E>
  __device__ __forceinline__ void rec_issue(double* sbuf, const double* grec, int bi) {
    if constexpr (BULK) {
      __syncwarp();  // every lane is done reading the buffer that is about to be overwritten
      if (lane == 0) bulk_load(sbuf, grec, E * TILE * 8, bars + bi);
    } else {
      stage_record<E>(sbuf, grec);
    }
  }
  __device__ __forceinline__ void rec_wait(int bi) {
    if constexpr (BULK) {
      mbar_wait(bars + bi, (bar_phase >> bi) & 1u);
      bar_phase ^= 1u << bi;
    }
  }
  // make this thread's earlier generic-proxy stores of records visible to later (async-proxy) bulk reads
  __device__ __forceinline__ void rec_fence() {
    __threadfence();
    if constexpr (BULK) asm volatile("fence.proxy.async;" ::: "memory");
  }
  // double-buffered record stream: issue the copy of cell `tn` (if valid) while cell `t` is consumed
  template <int E>
  __device__ __forceinline__ const double* stream(double* base_g, int Erec, int t, int tn, bool tn_valid) {
    if constexpr (LY::STAGED) {
      double* cur = stage + (t & 1) * (LY::E_STAGE_TOT * TILE);
      if (tn_valid) {
        double* nxt = stage + (tn & 1) * (LY::E_STAGE_TOT * TILE);
        rec_issue<E>(nxt, rec(base_g, tn, Erec), tn & 1);
        if (META && smeta()) stage_meta(nxt, tn);
      }
      if constexpr (!BULK) {
        stage_commit();
        stage_wait<1>();
      }
      rec_wait(t & 1);
      return cur;
    } else {
      return rec(base_g, t, Erec);
    }
  }
  template <int E>
  __device__ __forceinline__ void stream_begin(double* base_g, int Erec, int t) {
    rec_fence();  // records written by earlier sweeps of this thread are read back through cp.async / bulk copies
    if constexpr (LY::STAGED) {
      double* nxt = stage + (t & 1) * (LY::E_STAGE_TOT * TILE);
      rec_issue<E>(nxt, rec(base_g, t, Erec), t & 1);
      if (META && smeta()) stage_meta(nxt, t);
      if constexpr (!BULK) stage_commit();
    }
  }
  __device__ __forceinline__ void stream_end() {
    if constexpr (LY::STAGED && !BULK) stage_wait<0>();
  }
  __device__ __forceinline__ void load_zterm(double* zt) const {
    if (p.z_term_pp) {
      const double* q = p.z_term_pp + ((size_t)tile * DZT) * TILE + lane;
#pragma unroll
      for (int a = 0; a < DZT; ++a) zt[a] = q[a * TILE];
    } else {
#pragma unroll
      for (int a = 0; a < DZT; ++a) zt[a] = p.z_term[a];
    }
  }
  __device__ __forceinline__ double cell_alpha(int t, int flags, double alpha) const {
    if constexpr (HOT == 1) return alpha;
    // (HOT = 2: PLAIN cells pass flags = 0 and this folds to `alpha`; sweeps that contain a cell with its own alpha -- MPC: the
    // cell appended by the horizon shift -- take the generic-flag loop, see sweep_is_plain)
    return ((flags & I2C_CELL_OWN_ALPHA) && own_alpha_valid) ? p.alpha_cell[(size_t)slot(t) * p.Bpad + b] : alpha;
  }

  // Throughput variants: branch-free exp with constant-bank coefficients (+1.5 % at 65536 problems).  Latency variants keep
  // the library routine: measured, the custom one makes ptxas schedule the forward cell worse (-3 % at 4096 problems,
  // Estrin or Horner alike) although its dependency chain is shorter.  The two agree to <= 2 ulp.
  __device__ __forceinline__ static double pdf_exp(double x) {
#ifdef I2C_LAT_LIBEXP
    if constexpr (META) return exp(x);
#endif
    if constexpr (META) return fast_exp_neg_lat(x);
    else return fast_exp_neg(x);
  }
  // exp(-1/2 d^T C^-1 d): the pdf ratio w/Z of i2c.py:369-374 (scipy multivariate_normal)
  __device__ __forceinline__ bool pdf_ratio(const double* C_in, const double* d_in, double& rho) {
    if constexpr (DX == 2) {
      // closed form for 2x2 (one reciprocal instead of two dependent rsqrt pivots on the critical path)
      const double c00 = C_in[0], c10 = C_in[1], c11 = C_in[2];
      const double det = fma(c00, c11, -c10 * c10);
      const double q = fma(c11 * d_in[0], d_in[0], fma(-2.0 * c10 * d_in[0], d_in[1], c00 * d_in[1] * d_in[1])) * fast_rcp(det);
      rho = pdf_exp(-0.5 * q);
      return (c00 > 0.0) && (det > 0.0) && (det < kFm[20]);
    } else {
      double C[TRI(DX)], invd[DX], d[DX];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) C[i] = C_in[i];
#pragma unroll
      for (int i = 0; i < DX; ++i) d[i] = d_in[i];
      bool ok = chol_rows<DX>(C, invd);
      fwd_subst<DX>(C, invd, d);
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < DX; ++i) q = fma(d[i], d[i], q);
      rho = pdf_exp(-0.5 * q);
      return ok;
    }
  }

  // Joint (x,u) Gaussian under a linear-Gaussian controller around the incoming state message:
  //   mu = [m0; mu_u + Kt (m0 - mu_x_ref)],  Sigma = [[S0, S0 Kt^T],[Kt S0, Su]]
  // Writes mu[N], Sig[TRI N] and the Cholesky factor L[TRI N] (rows < DX reuse the carried factor).
  __device__ __forceinline__ bool build_joint(const Carry<DX>& c, const double* Kt, const double* mu_u, const double* Suu,
                                              bool coupled, double* mu, double* Sig, double* L, double* invd) {
#pragma unroll
    for (int i = 0; i < DX; ++i) mu[i] = c.m[i];
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) {
      Sig[i] = c.S[i];
      L[i] = c.L[i];
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) invd[i] = c.invd[i];
#pragma unroll
    for (int r = 0; r < DU; ++r) {
      mu[DX + r] = mu_u[r];
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        double s = 0.0;
        if (coupled) {
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], c.S[six(k, j)], s);
        }
        Sig[tix(DX + r, j)] = s;
        L[tix(DX + r, j)] = s;
      }
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        Sig[tix(DX + r, DX + q)] = Suu[tix(r, q)];
        L[tix(DX + r, DX + q)] = Suu[tix(r, q)];
      }
    }
    return chol_rows<N, DX>(L, invd);
  }

  // ---------------------------------------------------------------------------------- Linearize moments
  // First-order moments of the cost-feature map around the mean (observe_linearize, env_def.py:171-181, 278-298,
  // 541-570, 700-761): every feature is either the identity of component idx or sin / cos of the angle at idx, so
  // H = d z / d xu has one non-zero per row: z_a ~ h_a * xu[idx_a].  What _forward_msgs_linearize (i2c.py:297-306)
  // feeds its Kalman update:  mz = observe(mu), Sxy = Sigma H^T, Sz = H Sigma H^T (noise added by the caller).
  // CROSS = false drops the x-u cross terms: sig_z0_m = C sig_x C^T + D sig_u D^T of the backward pass (i2c.py:538-540).
  template <int D, int DY, bool TERM, bool CROSS>
  __device__ __forceinline__ void lin_obs_moments(const double* mu, const double* Sig, double* mz, double* Sz, double* Sxy) {
    int idx[DY];
    double h[DY];
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      const int sa = TERM ? Env::term_src(a) : Env::obs_src(a);
      if (sa >= 0) {
        idx[a] = sa;
        h[a] = 1.0;
        mz[a] = mu[sa];
      } else {
        const int k = -1 - sa, ang = Env::nl_angle(k);
        double sn, cs;
        sincos(mu[ang], &sn, &cs);
        idx[a] = ang;
        mz[a] = (k & 1) ? cs : sn;  // nonlinear features come in (sin, cos) pairs
        h[a] = (k & 1) ? -sn : cs;
      }
    }
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      if (Sxy) {
#pragma unroll
        for (int i = 0; i < D; ++i) Sxy[i * DY + a] = h[a] * Sig[six(i, idx[a])];
      }
#pragma unroll
      for (int bb = 0; bb <= a; ++bb) {
        const bool same_block = CROSS || ((idx[a] < DX) == (idx[bb] < DX));
        Sz[tix(a, bb)] = same_block ? h[a] * h[bb] * Sig[six(idx[a], idx[bb])] : 0.0;
      }
    }
  }
  // Linearised dynamics around mu (forward_linearize, model.py:158-164 / 240-242): m3 = f(mu), AB = df/dxu by
  // forward-mode AD on the in-kernel dynamics (the reference uses autograd), Sxy = Sigma AB^T, S3 = AB Sigma AB^T.
  __device__ __forceinline__ void lin_dyn_moments(const double* mu, const double* Sig, double* m3, double* S3, double* Sxy) {
    Dual<N> x[N], y[DX];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = dvar<N>(mu[i], i);
    Env::template dyn_g<Dual<N>>(x, par, y);
#pragma unroll
    for (int r = 0; r < DX; ++r) m3[r] = y[r].v;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int r = 0; r < DX; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(Sig[six(i, k)], y[r].d[k], s);
        Sxy[i * DX + r] = s;
      }
#pragma unroll
    for (int r = 0; r < DX; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s = fma(y[r].d[i], Sxy[i * DX + q], s);
        S3[tix(r, q)] = s;
      }
  }
  __device__ __forceinline__ static constexpr bool lin() { return LIN; }
  // sigma-point transform with the graph's rule: cubature points (sf, w0, wi) or the Gauss-Hermite grid
  template <int D_, int DY_, class Eval>
  __device__ __forceinline__ void xform(const double* m, const double* L, double sf, double w0, double wi, Eval&& eval,
                                        double* my, double* Syy, double* Dm) const {
    if constexpr (GH) grid_transform<D_, DY_>(m, L, p.gh, eval, my, Syy, Dm);
    else sigma_transform<D_, DY_, TrigT::kPM>(m, L, sf, HOT ? 0.0 : w0, wi, eval, my, Syy, Dm);  // HOT: zero centre weight
  }

  // ---------------------------------------------------------------------------------- forward cell
  // I2cCell._forward_msgs_quadrature (i2c.py:350-447); with p.linearize (linear envs) the same cell with exact
  // linear moments = _forward_msgs_linearize (i2c.py:244-348; the terminal update then happens in the backward pass).  c: (mu_x0_f, sig_x0_f) in, (mu_x3_f, sig_x3_f) out.
  // PLAIN: the cell is known to be a feedback (not independent), non-terminal cell -- the sweep checked the flags of the
  // whole horizon beforehand (sweep_is_plain) -- so neither branch exists in the loop body.
  // With PLAIN the trigonometric context of the cost-feature transform comes in through octx and is re-evaluated for the
  // NEXT cell at the end, from the outgoing message.  PS = element stride of the prior record at pr (1: register copy).
  template <bool PLAIN = false, int PS = TILE, bool JO = false>
  __device__ __forceinline__ void forward_cell(int it, int t, int flags, double alpha, bool aux, const double* pr,
                                               Carry<DX>& c, LogAcc& ent_x, TrigT* octx = nullptr, double* fr_in = nullptr) {
    double mu[N], Sig[TRI(N)], L[TRI(N)], invd[N];
    {
      double mu_u[DU], Suu[TRI(DU)], Kt[DU * DX];
#pragma unroll
      for (int r = 0; r < DU; ++r) mu_u[r] = pr[(LY::P_MU + DX + r) * PS];
#pragma unroll
      for (int r = 0; r < DU; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) Suu[tix(r, q)] = pr[(LY::P_SIG + tix(DX + r, DX + q)) * PS];
      const bool indep = PLAIN ? false : (flags & I2C_CELL_INDEPENDENT);
      if (!indep) {
        // feedback prior (i2c.py:361-387): K <- K * N(mu_x0_f; mu_prev, C)/N(mu_prev; mu_prev, C), C = Sig_xx + sig_x0_f
        double mx[DX], Sxx[TRI(DX)], Sux[DU * DX], C[TRI(DX)], d[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) mx[i] = pr[(LY::P_MU + i) * PS];
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) Sxx[i] = pr[(LY::P_SIG + i) * PS];
#pragma unroll
        for (int r = 0; r < DU; ++r)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            Sux[r * DX + j] = pr[(LY::P_SIG + tix(DX + r, j)) * PS];
            Kt[r * DX + j] = pr[(LY::P_K + r * DX + j) * PS];
          }
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) C[i] = Sxx[i] + c.S[i];
#pragma unroll
        for (int i = 0; i < DX; ++i) d[i] = c.m[i] - mx[i];
        double rho = 1.0;
#ifndef I2C_NO_RHO_LINEAR
        if constexpr (!LIN) {
          // Everything below is linear or quadratic in rho: form the rho-free parts first so that they overlap with the
          // exp() chain of the pdf ratio, and apply rho afterwards (2-3 dependent operations instead of ~15).
          double KS0[DU * DX], kd0[DU], A0[TRI(DU)], B0[TRI(DU)], W0[DU * DX];
#pragma unroll
          for (int r = 0; r < DU; ++r) {
            double sd = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) sd = fma(Kt[r * DX + k], d[k], sd);
            kd0[r] = sd;
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              double s = 0.0;
#pragma unroll
              for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], c.S[six(k, j)], s);
              KS0[r * DX + j] = s;
              W0[r * DX + j] = s;
            }
            fwd_subst<DX>(c.L, c.invd, W0 + r * DX);
          }
#pragma unroll
          for (int r = 0; r < DU; ++r)
#pragma unroll
            for (int q = 0; q <= r; ++q) {
              double a = 0.0, b = 0.0;
#pragma unroll
              for (int k = 0; k < DX; ++k) {
                a = fma(Kt[r * DX + k], Sux[q * DX + k], a);
                b = fma(KS0[r * DX + k], Kt[q * DX + k], b);
              }
              A0[tix(r, q)] = a;
              B0[tix(r, q)] = b;
            }
          if (!pdf_ratio(C, d, rho)) fail(I2C_FAIL_MVN, it, t);
          const double rho2 = rho * rho;
#pragma unroll
          for (int i = 0; i < DX; ++i) mu[i] = c.m[i];
#pragma unroll
          for (int i = 0; i < TRI(DX); ++i) {
            Sig[i] = c.S[i];
            L[i] = c.L[i];
          }
#pragma unroll
          for (int i = 0; i < DX; ++i) invd[i] = c.invd[i];
#pragma unroll
          for (int r = 0; r < DU; ++r) {
            mu[DX + r] = fma(rho, kd0[r], mu_u[r]);
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              Sig[tix(DX + r, j)] = rho * KS0[r * DX + j];
              L[tix(DX + r, j)] = rho * W0[r * DX + j];
            }
#pragma unroll
            for (int q = 0; q <= r; ++q) {
              const double v = fma(rho2, B0[tix(r, q)], fma(-rho, A0[tix(r, q)], Suu[tix(r, q)]));
              Sig[tix(DX + r, DX + q)] = v;
              L[tix(DX + r, DX + q)] = v;
            }
          }
          if (!chol_rows_pre<N, DX>(L, invd)) fail(I2C_FAIL_CHOL_PRIOR, it, t);
          goto joint_done;
        }
#endif
        // the quadrature cell always applies the ratio (quirk A.6.3); the linearize cell only for expert cells (:259)
        if (!lin() || (flags & I2C_CELL_EXPERT)) {
          if (!pdf_ratio(C, d, rho)) fail(I2C_FAIL_MVN, it, t);
        }
#pragma unroll
        for (int i = 0; i < DU * DX; ++i) Kt[i] *= rho;
        // mu_u0_f = mu_u0_m + K (mu_x0_f - mu_x0_m);  sig_u0_f = sig_u0_m - K sig_ux^T + K sig_x0_f K^T
        double KS[DU * DX];
#pragma unroll
        for (int r = 0; r < DU; ++r)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], c.S[six(k, j)], s);
            KS[r * DX + j] = s;
          }
#pragma unroll
        for (int r = 0; r < DU; ++r) {
          double s = mu_u[r];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], d[k], s);
          mu_u[r] = s;
#pragma unroll
          for (int q = 0; q <= r; ++q) {
            double v = Suu[tix(r, q)];
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(-Kt[r * DX + k], Sux[q * DX + k], v);
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(KS[r * DX + k], Kt[q * DX + k], v);
            Suu[tix(r, q)] = v;
          }
        }
      }
      if (!build_joint(c, Kt, mu_u, Suu, !indep, mu, Sig, L, invd)) fail(I2C_FAIL_CHOL_PRIOR, it, t);
    }
#ifndef I2C_NO_RHO_LINEAR
  joint_done:
#endif
    double* af = aux ? rec(p.auxf, t, LY::E_AUXF) : nullptr;
    if (aux) {
#pragma unroll
      for (int i = 0; i < N; ++i) af[(LY::AF_MU0 + i) * TILE] = mu[i];
#pragma unroll
      for (int i = 0; i < TRI(N); ++i) af[(LY::AF_SIG0 + i) * TILE] = Sig[i];
    }

    // ---- cost observation update (i2c.py:390-404)
    {
      double mz[DZ], Sz[TRI(DZ)], Sxy[N * DZ], z[DZ];
      if constexpr (LIN) {
        lin_obs_moments<N, DZ, false, true>(mu, Sig, mz, Sz, Sxy);
      } else if (fobs()) {
        structured_obs_moments<Env, N, DZ, false, TrigT>(mu, Sig, L, p.sf_n, p.wi_n, mz, Sz, Sxy, PLAIN ? octx : nullptr);
      } else {
        TrigT ctx;
        Env::center(mu, ctx);
        Env::offsets(L, p.sf_n, ctx);
        xform<N, DZ>(mu, L, p.sf_n, p.w0_n, p.wi_n,
                               [&](const double* x, int j, double* y) { Env::obs(x, j, ctx, y); }, mz, Sz, Sxy);
        cross_cov<N, DZ>(L, Sxy);
      }
      const double a_cell = cell_alpha(t, flags, alpha);
#pragma unroll
      for (int a = 0; a < DZ; ++a)
#pragma unroll
        for (int bb = 0; bb <= a; ++bb) Sz[tix(a, bb)] = fma(a_cell, p.QRinv[a * DZ + bb], Sz[tix(a, bb)]);
      if (aux) {
#pragma unroll
        for (int i = 0; i < DZ; ++i) af[(LY::AF_MUZ + i) * TILE] = mz[i];
#pragma unroll
        for (int i = 0; i < TRI(DZ); ++i) af[(LY::AF_SIGZ + i) * TILE] = Sz[i];
      }
      staged_z(pr, t, z);
      if (!condition<N, DZ>(mu, Sig, Sz, Sxy, mz, z)) fail(I2C_FAIL_CHOL_OBS, it, t);
    }
    double* fr;  // PLAIN: the caller's record cursor (compile-time choice: a run-time test would be a branch in the cell)
    if constexpr (PLAIN) fr = fr_in;
    else fr = rec(p.filt, t, LY::E_FILT);
#pragma unroll
    for (int i = 0; i < N; ++i) fr[(LY::F_MU1 + i) * TILE] = mu[i];
#pragma unroll
    for (int i = 0; i < TRI(N); ++i) fr[(LY::F_SIG1 + i) * TILE] = Sig[i];

    // ---- dynamics moment matching (i2c.py:415-428)
#pragma unroll
    for (int i = 0; i < TRI(N); ++i) L[i] = Sig[i];
    if (!chol_rows<N>(L, invd)) fail(I2C_FAIL_CHOL_FILTERED, it, t);
    {
      double Sxy[N * DX];
      if constexpr (LIN) {
        lin_dyn_moments(mu, Sig, c.m, c.S, Sxy);
      } else {
        TrigT ctx;
        Env::center(mu, ctx);
        Env::offsets(L, p.sf_n, ctx);
        xform<N, DX>(mu, L, p.sf_n, p.w0_n, p.wi_n,
                               [&](const double* x, int j, double* y) { Env::dyn(x, j, ctx, par, y); }, c.m, c.S, Sxy);
        cross_cov<N, DX>(L, Sxy);
      }
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) {
        c.S[i] += p.sig_eta[i];
        c.L[i] = c.S[i];
      }
      if (!chol_rows<DX>(c.L, c.invd)) fail(I2C_FAIL_CHOL_X3, it, t);
      // J_dyn = Sxy Sig_x3^{-1}  (n x dx)
      if constexpr (JO) {
        // hand S_xy and the factor over to helper warp 1 (jdyn_serve); slot reuse two cells later: the helper is at most
        // one cell behind (checked, never waited for in practice: a J job is ~5 % of a cell)
        const unsigned sl = jcount & 1u, par = (jcount >> 1) & 1u;
        mbar_wait_s(bars_s + 8u * (JBAR + 2 + sl), par ^ 1u);  // passes at once for the first use of a slot
        const unsigned js = jslot_s + sl * (JELEMS * TILE * 8);
#pragma unroll
        for (int i = 0; i < N * DX; ++i) sts_f64(js + i * (TILE * 8), Sxy[i]);
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) sts_f64(js + (N * DX + i) * (TILE * 8), c.L[i]);
#pragma unroll
        for (int i = 0; i < DX; ++i) sts_f64(js + (N * DX + TRI(DX) + i) * (TILE * 8), c.invd[i]);
        mbar_arrive_s(bars_s + 8u * (JBAR + sl));
        ++jcount;
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double w[DX];
#pragma unroll
          for (int a = 0; a < DX; ++a) w[a] = Sxy[i * DX + a];
          fwd_subst<DX>(c.L, c.invd, w);
          bwd_subst<DX>(c.L, c.invd, w);
#pragma unroll
          for (int a = 0; a < DX; ++a) fr[(LY::F_J + i * DX + a) * TILE] = w[a];
        }
      }
    }
    // ---- terminal cost update on the outgoing message (i2c.py:430-443)
    if (!PLAIN && Env::HAS_TERM && (flags & I2C_CELL_TERMINAL) && p.has_qf && !lin()) {
      double mz[DZT], Sz[TRI(DZT)], Sxy[DX * DZT];
      if (fobs()) {
        structured_obs_moments<Env, DX, DZT, true, TrigT>(c.m, c.S, c.L, p.sf_x, p.wi_x, mz, Sz, Sxy);
      } else {
        TrigT ctx;
        Env::center(c.m, ctx);
        Env::offsets(c.L, p.sf_x, ctx);
        xform<DX, DZT>(c.m, c.L, p.sf_x, p.w0_x, p.wi_x,
                                 [&](const double* x, int j, double* y) { Env::obs_term(x, j, ctx, y); }, mz, Sz, Sxy);
        cross_cov<DX, DZT>(c.L, Sxy);
      }
      const double a_cell = cell_alpha(t, flags, alpha);
#pragma unroll
      for (int a = 0; a < DZT; ++a)
#pragma unroll
        for (int bb = 0; bb <= a; ++bb) Sz[tix(a, bb)] = fma(a_cell, p.Qfinv[a * DZT + bb], Sz[tix(a, bb)]);
      double zt[DZT];
      load_zterm(zt);
      if (!condition<DX, DZT>(c.m, c.S, Sz, Sxy, mz, zt)) fail(I2C_FAIL_CHOL_TERMINAL, it, t);
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) c.L[i] = c.S[i];
      if (!chol_rows<DX>(c.L, c.invd)) fail(I2C_FAIL_CHOL_TERMINAL, it, t);
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      fr[(LY::F_MU3 + i) * TILE] = c.m[i];
      ent_x.mul(c.L[tix(i, i)]);
    }
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) fr[(LY::F_SIG3 + i) * TILE] = c.S[i];
    if constexpr (PLAIN && Env::OBS_NL > 0) obs_trig(c, *octx);
  }
  // trigonometric context of the cost-feature transform of the cell that receives message c: the angles are state
  // components, and the state block of the joint's mean / factor is the message itself
  __device__ __forceinline__ void obs_trig(const Carry<DX>& c, TrigT& o) const {
    Env::center(c.m, o);
    Env::offsets(c.L, p.sf_n, o);
  }

  // helper warp 1, during the forward sweep: the J_dyn rows of the T cells warp 0 hands over (see JOFF)
  __device__ __forceinline__ void jdyn_serve(int T) {
    for (int t = 0; t < T; ++t) {
      const unsigned sl = jcount & 1u, par = (jcount >> 1) & 1u;
      while (!mbar_test_s(bars_s + 8u * (JBAR + sl), par)) __nanosleep(500);  // a J job is not urgent: poll rarely
      const unsigned js = jslot_s + sl * (JELEMS * TILE * 8);
      double Lx[TRI(DX)], ix[DX];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) Lx[i] = lds_f64(js + (N * DX + i) * (TILE * 8));
#pragma unroll
      for (int i = 0; i < DX; ++i) ix[i] = lds_f64(js + (N * DX + TRI(DX) + i) * (TILE * 8));
      double* fr = rec(p.filt, t, LY::E_FILT);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double w[DX];
#pragma unroll
        for (int a = 0; a < DX; ++a) w[a] = lds_f64(js + (i * DX + a) * (TILE * 8));
        if (i == N - 1) mbar_arrive_s(bars_s + 8u * (JBAR + 2 + sl));  // the slot is in registers
        fwd_subst<DX>(Lx, ix, w);
        bwd_subst<DX>(Lx, ix, w);
#pragma unroll
        for (int a = 0; a < DX; ++a) fr[(LY::F_J + i * DX + a) * TILE] = w[a];
      }
      ++jcount;
    }
    // the J rows are read back by the copy warp's bulk copies after the block-wide barrier that ends the forward sweep
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  }

  // ---------------------------------------------------------------------------------- backward cell
  // I2cCell._backward_msgs_quadrature (i2c.py:544-610) for a non-final cell, with the per-cell M-step
  // statistics fused in.  (m3m, S3m) in: next cell's (mu_x0_m, sig_x0_m); out: this cell's.
  struct Stats {
    double cost, cost_var, tr;
    LogAcc ent_u;
  };
  // RTS recursion of one cell (i2c.py:578-592): posterior joint from the filtered record and the next cell's
  // smoothed state; stores mu_xu0_m / sig_xu0_m and hands (mu_x0_m, sig_x0_m) to the previous cell.
  // FS = element stride of the filtered record at fr: TILE for the tiled layouts in shared / global memory, 1 for a copy
  // held in registers
  template <int FS = TILE, bool PO = false>
  __device__ __forceinline__ void backward_head(int it, int t, bool aux, const double* fr, double* m3m, double* S3m,
                                                double* mu, double* Sig, double* po_in = nullptr) {
    double J[N * DX];
    {
      double dm[DX], dS[TRI(DX)];
#pragma unroll
      for (int i = 0; i < DX; ++i) dm[i] = m3m[i] - fr[(LY::F_MU3 + i) * FS];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) dS[i] = S3m[i] - fr[(LY::F_SIG3 + i) * FS];
#pragma unroll
      for (int i = 0; i < N * DX; ++i) J[i] = fr[(LY::F_J + i) * FS];
      if (aux) {
        double* ab = rec(p.auxb, t, LY::E_AUXB);
#pragma unroll
        for (int i = 0; i < DX; ++i) ab[(LY::AB_MU3M + i) * TILE] = m3m[i];
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) ab[(LY::AB_SIG3M + i) * TILE] = S3m[i];
      }
      // mu_xu1_m = mu_xu1_f + J (mu_x3_m - mu_x3_f);  sig_xu1_m = sig_xu1_f + J (sig_x3_m - sig_x3_f) J^T
      double JD[N * DX];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = fr[(LY::F_MU1 + i) * FS];
#pragma unroll
        for (int k = 0; k < DX; ++k) s = fma(J[i * DX + k], dm[k], s);
        mu[i] = s;
#pragma unroll
        for (int k = 0; k < DX; ++k) {
          double v = 0.0;
#pragma unroll
          for (int l = 0; l < DX; ++l) v = fma(J[i * DX + l], dS[six(l, k)], v);
          JD[i * DX + k] = v;
        }
      }
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double s = fr[(LY::F_SIG1 + tix(i, j)) * FS];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(JD[i * DX + k], J[j * DX + k], s);
          Sig[tix(i, j)] = s;
        }
    }
    double* po;
    if constexpr (PO) po = po_in;
    else po = rec(post, t, LY::E_POST);
#pragma unroll
    for (int i = 0; i < N; ++i) po[(LY::P_MU + i) * TILE] = mu[i];
#pragma unroll
    for (int i = 0; i < TRI(N); ++i) po[(LY::P_SIG + i) * TILE] = Sig[i];
#pragma unroll
    for (int i = 0; i < DX; ++i) m3m[i] = mu[i];
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) S3m[i] = Sig[i];

  }

  // Everything of the backward cell that does NOT feed the recursion (i2c.py:594-608 + M-step statistics):
  // Cholesky of the posterior joint, controller K / k / sigK, marginal cost-feature moments, cost / alpha / entropy
  // statistics.  Independent across cells => can run on other warps (team kernel).  zbuf: staged targets or NULL.
  __device__ __forceinline__ void backward_tail(int it, int t, bool aux, const double* zbuf, const double* mu,
                                                const double* Sig, Stats& st) {
    double* po = rec(post, t, LY::E_POST);
    double L[TRI(N)], invd[N];
#pragma unroll
    for (int i = 0; i < TRI(N); ++i) L[i] = Sig[i];
    if (!chol_rows<N>(L, invd)) fail(I2C_FAIL_CHOL_POSTERIOR, it, t);

    // controller (i2c.py:598-608): K = sig_ux sig_xx^{-1} = L_ux L_xx^{-1}; sigK = sig_uu - K sig_ux^T = L_uu L_uu^T
#pragma unroll
    for (int r = 0; r < DU; ++r) {
      double w[DX];
#pragma unroll
      for (int j = 0; j < DX; ++j) w[j] = L[tix(DX + r, j)];
      bwd_subst<DX>(L, invd, w);  // solves L_xx^T k^T = l^T  <=>  k L_xx = l
      double kk = mu[DX + r];
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        po[(LY::P_K + r * DX + j) * TILE] = w[j];
        kk = fma(-w[j], mu[j], kk);
      }
      po[(LY::P_KK + r) * TILE] = kk;
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k <= q; ++k) s = fma(L[tix(DX + r, DX + k)], L[tix(DX + q, DX + k)], s);
        po[(LY::P_SIGK + tix(r, q)) * TILE] = s;
      }
    }
    // HOT = 2 (closed-loop MPC sweeps: forward + backward + _update_priors, policy/mpc.py:147-154): without the M-step phase
    // nobody reads the entropy / cost / alpha statistics below -- the tail is the controller extraction only
    bool stats = true;
    if constexpr (HOT == 2) stats = (p.phases & I2C_PH_MSTEP) != 0;
    if (!stats) return;
    // policy entropy needs det(sig_u0_m) (i2c.py:1072-1081)
    {
      double Su[TRI(DU)], iu[DU];
#pragma unroll
      for (int r = 0; r < DU; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) Su[tix(r, q)] = Sig[tix(DX + r, DX + q)];
      if (!chol_rows<DU>(Su, iu)) fail(I2C_FAIL_POLICY_DET, it, t);
#pragma unroll
      for (int r = 0; r < DU; ++r) st.ent_u.mul(Su[tix(r, r)]);
    }
    // marginal cost-feature moments (i2c.py:594-596) -> alpha / cost statistics
    {
      double mz[DZ], Sz[TRI(DZ)], z[DZ];
      if constexpr (LIN) {
        // mu_z0_m = observe(mu_xu0_m); sig_z0_m = C sig_x0_m C^T + D sig_u0_m D^T (no cross term, i2c.py:538-540)
        lin_obs_moments<N, DZ, false, false>(mu, Sig, mz, Sz, nullptr);
      } else if (fobs()) {
        double Cxy[N * DZ];
        structured_obs_moments<Env, N, DZ, false, TrigT>(mu, Sig, L, p.sf_n, p.wi_n, mz, Sz, Cxy);
      } else {
        double Dm[N * DZ];
        TrigT ctx;
        Env::center(mu, ctx);
        Env::offsets(L, p.sf_n, ctx);
        xform<N, DZ>(mu, L, p.sf_n, p.w0_n, p.wi_n,
                               [&](const double* x, int j, double* y) { Env::obs(x, j, ctx, y); }, mz, Sz, Dm);
      }
      if (aux) {
        double* ab = rec(p.auxb, t, LY::E_AUXB);
#pragma unroll
        for (int i = 0; i < DZ; ++i) ab[(LY::AB_MUZ + i) * TILE] = mz[i];
#pragma unroll
        for (int i = 0; i < TRI(DZ); ++i) ab[(LY::AB_SIGZ + i) * TILE] = Sz[i];
      }
      double cm, cv;
      if constexpr (LIN) {
        // calc_cost goes through the graph's cubature transform of the joint posterior (i2c.py:1034-1043), which
        // is exact for a linear map: full E F Sigma (E F)^T including the x-u cross terms
        // (for the nonlinear envs the graph's cubature transform is used, as in the reference)
        double mzf[DZ], Szf[TRI(DZ)], Sxyf[N * DZ];
        if (Env::OBS_NL == 0) {
          lin_obs_moments<N, DZ, false, true>(mu, Sig, mzf, Szf, Sxyf);
        } else {
          structured_obs_moments<Env, N, DZ, false, TrigT>(mu, Sig, L, p.sf_n, p.wi_n, mzf, Szf, Sxyf);
        }
        cost_stats<DZ>(p, mzf, Szf, p.z_graph, cm, cv);
      } else {
        cost_stats<DZ>(p, mz, Sz, p.z_graph, cm, cv);
      }
      st.cost += cm;
      st.cost_var += cv;
      if (HOT || zbuf) staged_z(zbuf, t, z); else load_z(t, z);  // HOT: the shared-memory table
      st.tr += alpha_trace<DZ>(p.QR, p.qr_diag, mz, Sz, z);
    }
  }

  __device__ __forceinline__ void backward_cell(int it, int t, int flags, bool aux, const double* fr, double* m3m,
                                                double* S3m, Stats& st) {
    double mu[N], Sig[TRI(N)];
    backward_head(it, t, aux, fr, m3m, S3m, mu, Sig);
    backward_tail(it, t, aux, fr, mu, Sig, st);
  }

  // end-of-chain handling of the last cell (i2c.py:546-572): covariance control or plain hand-over, and
  // the terminal cost-feature moments for the alpha update.  c holds (mu_x3_f, sig_x3_f, chol).
  __device__ __forceinline__ void backward_terminal(int it, int t, double temp, double a_cell, const Carry<DX>& c,
                                                    double* m3m, double* S3m, double& tr_term) {
    if constexpr (LIN) {
      // _backward_msgs_linearize, end of chain (i2c.py:450-501)
      tr_term = 0.0;
#pragma unroll
      for (int i = 0; i < DX; ++i) m3m[i] = c.m[i];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) S3m[i] = c.S[i];
      if (p.cov_ctrl) {
#pragma unroll
        for (int i = 0; i < DX; ++i) m3m[i] = p.mu_xt[i];
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) S3m[i] = p.sxt[i];
      } else if (Env::HAS_TERM && p.has_qf) {
        // terminal cost as a Kalman update through observe_terminal_linearize at mu_x3_f (i2c.py:474-492)
        double mz[DZT], Sz[TRI(DZT)], Sxy[DX * DZT], zt[DZT];
        lin_obs_moments<DX, DZT, true, true>(c.m, c.S, mz, Sz, Sxy);
#pragma unroll
        for (int i = 0; i < DZT; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) Sz[tix(i, j)] = fma(a_cell, p.Qfinv[i * DZT + j], Sz[tix(i, j)]);
        load_zterm(zt);
        if (!condition<DX, DZT>(m3m, S3m, Sz, Sxy, mz, zt)) fail(I2C_FAIL_CHOL_TERMINAL, it, t);
      }
      if (Env::HAS_TERM && p.has_qf) {
        // mu_z3_m = observe_terminal(mu_x3_m); sig_z3_m = E sig_x3_m E^T + sig_xi_terminal (i2c.py:500-501)
        double mz3[DZT], Sz3[TRI(DZT)], zt[DZT];
        lin_obs_moments<DX, DZT, true, true>(m3m, S3m, mz3, Sz3, nullptr);
#pragma unroll
        for (int i = 0; i < DZT; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) Sz3[tix(i, j)] = fma(a_cell, p.Qfinv[i * DZT + j], Sz3[tix(i, j)]);
        double* tm = p.term + ((size_t)tile * LY::E_TERM) * TILE + lane;
#pragma unroll
        for (int i = 0; i < DZT; ++i) tm[(LY::TM_MU + i) * TILE] = mz3[i];
#pragma unroll
        for (int i = 0; i < TRI(DZT); ++i) tm[(LY::TM_SIG + i) * TILE] = Sz3[i];
        load_zterm(zt);
        tr_term = alpha_trace<DZT>(p.Qf, 0, mz3, Sz3, zt);
      }
      return;
    }
    double Lm[TRI(DX)], invm[DX];
    if (p.cov_ctrl) {
      // sig_x3_m = S - S (Sig_T + S)^{-1} S,  mu_x3_m = sig_x3_m (S^{-1} mu_x3_f + Sig_T^{-1} mu_T),  S = temp * sig_x3_f
      double S[TRI(DX)], A[TRI(DX)], ia[DX];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) {
        S[i] = temp * c.S[i];
        A[i] = p.sxt[i] + S[i];
      }
      if (!chol_rows<DX>(A, ia)) fail(I2C_FAIL_COV_CONTROL, it, t);
      double W[DX * DX];  // W[:, j] = La^{-1} S[:, j]
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        double w[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) w[i] = S[six(i, j)];
        fwd_subst<DX>(A, ia, w);
#pragma unroll
        for (int i = 0; i < DX; ++i) W[i * DX + j] = w[i];
      }
#pragma unroll
      for (int i = 0; i < DX; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double s = S[tix(i, j)];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(-W[k * DX + i], W[k * DX + j], s);
          S3m[tix(i, j)] = s;
        }
      // S^{-1} mu_x3_f = (1/temp) sig_x3_f^{-1} mu_x3_f via the carried factor
      double v[DX];
#pragma unroll
      for (int i = 0; i < DX; ++i) v[i] = c.m[i];
      fwd_subst<DX>(c.L, c.invd, v);
      bwd_subst<DX>(c.L, c.invd, v);
      const double it_ = 1.0 / temp;
#pragma unroll
      for (int i = 0; i < DX; ++i) v[i] = fma(v[i], it_, p.sxt_inv_mu[i]);
#pragma unroll
      for (int i = 0; i < DX; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DX; ++k) s = fma(S3m[six(i, k)], v[k], s);
        m3m[i] = s;
      }
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) Lm[i] = S3m[i];
      if (Env::HAS_TERM && p.has_qf) {
        if (!chol_rows<DX>(Lm, invm)) fail(I2C_FAIL_COV_CONTROL, it, t);
      }
    } else {
#pragma unroll
      for (int i = 0; i < DX; ++i) m3m[i] = c.m[i];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) {
        S3m[i] = c.S[i];
        Lm[i] = c.L[i];
      }
    }
    tr_term = 0.0;
    if (Env::HAS_TERM && p.has_qf) {
      double mz[DZT], Sz[TRI(DZT)], Dm[DX * DZT];
      TrigT ctx;
      Env::center(m3m, ctx);
      Env::offsets(Lm, p.sf_x, ctx);
      xform<DX, DZT>(m3m, Lm, p.sf_x, p.w0_x, p.wi_x,
                               [&](const double* x, int j, double* y) { Env::obs_term(x, j, ctx, y); }, mz, Sz, Dm);
      double* tm = p.term + ((size_t)tile * LY::E_TERM) * TILE + lane;
#pragma unroll
      for (int i = 0; i < DZT; ++i) tm[(LY::TM_MU + i) * TILE] = mz[i];
#pragma unroll
      for (int i = 0; i < TRI(DZT); ++i) tm[(LY::TM_SIG + i) * TILE] = Sz[i];
      double zt[DZT];
      load_zterm(zt);
      tr_term = alpha_trace<DZT>(p.Qf, 0, mz, Sz, zt);
    }
  }

  // ---------------------------------------------------------------------------------- propagate cell
  // I2cCell._propagate_forward_quadrature (i2c.py:150-199).
  struct PStats {
    double cost, cost_var, cost_min, tr;
    LogAcc ent;
  };
  template <int PS = TILE>
  __device__ __forceinline__ void propagate_cell(int it, int t, int flags, bool aux, const double* po, Carry<DX>& c,
                                                 PStats& st) {
    double mu[N], Sig[TRI(N)], L[TRI(N)], invd[N];
    {
      double mu_u[DU], Suu[TRI(DU)], Kt[DU * DX];
#pragma unroll
      for (int r = 0; r < DU; ++r) mu_u[r] = po[(LY::P_MU + DX + r) * PS];
#pragma unroll
      for (int r = 0; r < DU; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) Suu[tix(r, q)] = po[(LY::P_SIG + tix(DX + r, DX + q)) * PS];
#pragma unroll
      for (int i = 0; i < DU * DX; ++i) Kt[i] = po[(LY::P_K + i) * PS];
      if (!(flags & I2C_CELL_INDEPENDENT)) {
        double mx[DX], Sxx[TRI(DX)], d[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) mx[i] = po[(LY::P_MU + i) * PS];
#pragma unroll
        for (int i = 0; i < TRI(DX); ++i) Sxx[i] = po[(LY::P_SIG + i) * PS];
#pragma unroll
        for (int i = 0; i < DX; ++i) d[i] = c.m[i] - mx[i];
        if (flags & I2C_CELL_EXPERT) {
          double C[TRI(DX)], rho;
#pragma unroll
          for (int i = 0; i < TRI(DX); ++i) C[i] = Sxx[i] + c.S[i];
          // the reference swallows exceptions here (try/except + logging.error, i2c.py:161-167): K unchanged
          if (pdf_ratio(C, d, rho)) {
#pragma unroll
            for (int i = 0; i < DU * DX; ++i) Kt[i] *= rho;
          }
        }
        // mu_u0_pf = mu_u0_m + K (mu_x0_pf - mu_x0_m); sig_u0_pf = K sig_x0_pf K^T + sig_u0_m - K sig_x0_m K^T
#pragma unroll
        for (int r = 0; r < DU; ++r) {
          double s = mu_u[r];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], d[k], s);
          mu_u[r] = s;
        }
        double KD[DU * DX];  // K (sig_x0_pf - sig_x0_m)
#pragma unroll
        for (int r = 0; r < DU; ++r)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(Kt[r * DX + k], c.S[six(k, j)] - Sxx[six(k, j)], s);
            KD[r * DX + j] = s;
          }
#pragma unroll
        for (int r = 0; r < DU; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) {
            double v = Suu[tix(r, q)];
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(KD[r * DX + k], Kt[q * DX + k], v);
            Suu[tix(r, q)] = v;
          }
      }
      // NOTE: even for independent cells the joint carries the cross-covariance K sig_x (i2c.py:173-179)
      if (!build_joint(c, Kt, mu_u, Suu, true, mu, Sig, L, invd)) fail(I2C_FAIL_CHOL_PROPAGATE, it, t);
    }
    double* pf = aux ? rec(p.pf, t, LY::E_PF) : nullptr;
    if (aux) {
#pragma unroll
      for (int i = 0; i < N; ++i) pf[(LY::PF_MU + i) * TILE] = mu[i];
#pragma unroll
      for (int i = 0; i < TRI(N); ++i) pf[(LY::PF_SIG + i) * TILE] = Sig[i];
    }
    {
      double mz[DZ], Sz[TRI(DZ)], Dm[N * DZ], z[DZ];
      if (fobs()) {
        structured_obs_moments<Env, N, DZ, false, TrigT>(mu, Sig, L, p.sf_n, p.wi_n, mz, Sz, Dm);
      } else {
        TrigT ctx;
        Env::center(mu, ctx);
        Env::offsets(L, p.sf_n, ctx);
        xform<N, DZ>(mu, L, p.sf_n, p.w0_n, p.wi_n,
                               [&](const double* x, int j, double* y) { Env::obs(x, j, ctx, y); }, mz, Sz, Dm);
      }
      if (aux) {
#pragma unroll
        for (int i = 0; i < DZ; ++i) pf[(LY::PF_MUZ + i) * TILE] = mz[i];
#pragma unroll
        for (int i = 0; i < TRI(DZ); ++i) pf[(LY::PF_SIGZ + i) * TILE] = Sz[i];
      }
      double cm, cv;
      cost_stats<DZ>(p, mz, Sz, p.z_graph, cm, cv);
      st.cost += cm;
      st.cost_var += cv;
      st.cost_min = fmin(st.cost_min, cm);
      staged_z(po, t, z);
      st.tr += alpha_trace<DZ>(p.QR, p.qr_diag, mz, Sz, z);
    }
    {
      double Dm[N * DX];
      TrigT ctx;
      Env::center(mu, ctx);
      Env::offsets(L, p.sf_n, ctx);
      xform<N, DX>(mu, L, p.sf_n, p.w0_n, p.wi_n,
                             [&](const double* x, int j, double* y) { Env::dyn(x, j, ctx, par, y); }, c.m, c.S, Dm);
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) {
        c.S[i] += p.sig_eta[i];
        c.L[i] = c.S[i];
      }
      if (!chol_rows<DX>(c.L, c.invd)) fail(I2C_FAIL_CHOL_PROPAGATE, it, t);
#pragma unroll
      for (int i = 0; i < DX; ++i) st.ent.mul(c.L[tix(i, i)]);
    }
    if (aux) {
#pragma unroll
      for (int i = 0; i < DX; ++i) pf[(LY::PF_MU3 + i) * TILE] = c.m[i];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) pf[(LY::PF_SIG3 + i) * TILE] = c.S[i];
    }
  }

  // ---------------------------------------------------------------------------------- Riccati messages
  // I2cCell._backward_ricatti_msgs / I2cGraph._backward_ricatti_msgs (i2c.py:612-678, 888-893): backward
  // information-form recursion that reproduces the LQR value function; overwrites K, k, sigK of every cell.
  // Forward-pass quantities are re-derived from the stored records (prior joint = auxf, filtered, posterior).
  __device__ void riccati_sweep(double alpha) {
    if constexpr (Env::LINEAR) {
      constexpr int X2 = DX * DX;
      double Lb[X2], nb[DX];
      double A[X2], Bm[DX * DU], av[DX], Se[X2];
#pragma unroll
      for (int i = 0; i < X2; ++i) A[i] = par[i];
#pragma unroll
      for (int i = 0; i < DX * DU; ++i) Bm[i] = par[X2 + i];
#pragma unroll
      for (int i = 0; i < DX; ++i) av[i] = par[X2 + DX * DU + i];
#pragma unroll
      for (int i = 0; i < DX; ++i)
#pragma unroll
        for (int j = 0; j < DX; ++j) Se[i * DX + j] = p.sig_eta[six(i, j)];
      for (int t = p.T - 1; t >= 0; --t) {
        const double* af = rec(p.auxf, t, LY::E_AUXF);
        const double* fr = rec(p.filt, t, LY::E_FILT);
        double* po = rec(latest, t, LY::E_POST);
        const int flags = p.cell_flags[slot(t)];
        const double a_cell = cell_alpha(t, flags, alpha);
        double mu0[N], S0[TRI(N)], S1[TRI(N)], mu1[N], z[DZ];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          mu0[i] = af[(LY::AF_MU0 + i) * TILE];
          mu1[i] = fr[(LY::F_MU1 + i) * TILE];
        }
#pragma unroll
        for (int i = 0; i < TRI(N); ++i) {
          S0[i] = af[(LY::AF_SIG0 + i) * TILE];
          S1[i] = fr[(LY::F_SIG1 + i) * TILE];
        }
        load_z(t, z);
        if (t == p.T - 1) {
          // end of chain: nu_x3_b = sig_x3_m^{-1} mu_x3_m - nu_x3_f;  lambda_x3_b = sig_x3_m^{-1} - lambda_x3_f
          const double* ab = rec(p.auxb, t, LY::E_AUXB);
          double Sm[X2], Sf[X2], mm_[DX], mf[DX];
#pragma unroll
          for (int i = 0; i < DX; ++i) {
            mm_[i] = ab[(LY::AB_MU3M + i) * TILE];
            mf[i] = fr[(LY::F_MU3 + i) * TILE];
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              Sm[i * DX + j] = ab[(LY::AB_SIG3M + six(i, j)) * TILE];
              Sf[i * DX + j] = fr[(LY::F_SIG3 + six(i, j)) * TILE];
            }
          }
          inv_gj<DX>(Sm);
          inv_gj<DX>(Sf);
#pragma unroll
          for (int i = 0; i < DX; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              s = fma(Sm[i * DX + j], mm_[j], s);
              s = fma(-Sf[i * DX + j], mf[j], s);
              Lb[i * DX + j] = Sm[i * DX + j] - Sf[i * DX + j];
            }
            nb[i] = s;
          }
        }
        double* rr = rec(p.ric, t, LY::E_RIC);
#pragma unroll
        for (int i = 0; i < X2; ++i) rr[(LY::R_L3 + i) * TILE] = Lb[i];
#pragma unroll
        for (int i = 0; i < DX; ++i) rr[(LY::R_N3 + i) * TILE] = nb[i];
        // lambda_z1_f = (sig_xi + F sig_u0_f F^T)^{-1};  nu_z1_f = E^T lambda_z1 (z - F mu_u0_f)
        double Lz1[DZ * DZ], Lz2[DZ * DZ];
#pragma unroll
        for (int a = 0; a < DZ; ++a)
#pragma unroll
          for (int bb = 0; bb < DZ; ++bb) {
            double s1 = a_cell * p.QRinv[a * DZ + bb], s2 = s1;
#pragma unroll
            for (int i = 0; i < DU; ++i)
#pragma unroll
              for (int k = 0; k < DU; ++k) s1 = fma(Env::obsF(a, i) * Env::obsF(bb, k), S0[six(DX + i, DX + k)], s1);
#pragma unroll
            for (int i = 0; i < DX; ++i)
#pragma unroll
              for (int k = 0; k < DX; ++k) s2 = fma(Env::obsE(a, i) * Env::obsE(bb, k), S0[six(i, k)], s2);
            Lz1[a * DZ + bb] = s1;
            Lz2[a * DZ + bb] = s2;
          }
        inv_gj<DZ>(Lz1);
        inv_gj<DZ>(Lz2);
        double r1[DZ], r2[DZ], nz1[DX], Rug[DU], Q[X2];
#pragma unroll
        for (int a = 0; a < DZ; ++a) {
          double s1 = z[a], s2 = z[a];
#pragma unroll
          for (int i = 0; i < DU; ++i) s1 = fma(-Env::obsF(a, i), mu0[DX + i], s1);
#pragma unroll
          for (int i = 0; i < DX; ++i) s2 = fma(-Env::obsE(a, i), mu0[i], s2);
          r1[a] = s1;
          r2[a] = s2;
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DZ; ++a)
#pragma unroll
            for (int bb = 0; bb < DZ; ++bb) s = fma(Env::obsE(a, i) * Lz1[a * DZ + bb], r1[bb], s);
          nz1[i] = s;
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double q = 0.0;
#pragma unroll
            for (int a = 0; a < DZ; ++a)
#pragma unroll
              for (int bb = 0; bb < DZ; ++bb) q = fma(Env::obsE(a, i) * Lz1[a * DZ + bb], Env::obsE(bb, j), q);
            Q[i * DX + j] = q;
          }
        }
#pragma unroll
        for (int i = 0; i < DU; ++i) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DZ; ++a)
#pragma unroll
            for (int bb = 0; bb < DZ; ++bb) s = fma(Env::obsF(a, i) * Lz2[a * DZ + bb], r2[bb], s);
          Rug[i] = s;
        }
        // nu_u_0 = sig_u0_f^{-1} mu_u0_f
        double Su0[DU * DU], nu_u0[DU];
#pragma unroll
        for (int i = 0; i < DU; ++i)
#pragma unroll
          for (int k = 0; k < DU; ++k) Su0[i * DU + k] = S0[six(DX + i, DX + k)];
        inv_gj<DU>(Su0);
#pragma unroll
        for (int i = 0; i < DU; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DU; ++k) s = fma(Su0[i * DU + k], mu0[DX + k], s);
          nu_u0[i] = s;
        }
        // sig_u2_f = B sig_u1_f B^T; sig_x2_f = A sig_x1_f A^T + sig_eta; lambda_x2_f = inv(sig_x2_f)
        double Su2[X2], Sx2[X2], Lx2[X2], Sx1[X2], T1[X2];
#pragma unroll
        for (int i = 0; i < DX; ++i)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DU; ++k)
#pragma unroll
              for (int l = 0; l < DU; ++l) s = fma(Bm[i * DU + k] * S1[six(DX + k, DX + l)], Bm[j * DU + l], s);
            Su2[i * DX + j] = s;
            Sx1[i * DX + j] = S1[six(i, j)];
          }
        mm<DX, DX, DX>(A, Sx1, T1);
#pragma unroll
        for (int i = 0; i < DX; ++i)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = Se[i * DX + j];
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(T1[i * DX + k], A[j * DX + k], s);
            Sx2[i * DX + j] = s;
            Lx2[i * DX + j] = s;
          }
        inv_gj<DX>(Lx2);
        // gamma = lambda_x2 inv(lambda_x2 + Lb)
        double G1[X2], gamma[X2];
#pragma unroll
        for (int i = 0; i < X2; ++i) G1[i] = Lx2[i] + Lb[i];
        inv_gj<DX>(G1);
        mm<DX, DX, DX>(Lx2, G1, gamma);
        // M = inv(sig_eta + sig_u2) + Lb
        double M[X2], Minv[X2];
#pragma unroll
        for (int i = 0; i < X2; ++i) M[i] = Se[i] + Su2[i];
        inv_gj<DX>(M);
#pragma unroll
        for (int i = 0; i < X2; ++i) {
          M[i] += Lb[i];
          Minv[i] = M[i];
        }
        inv_gj<DX>(Minv);
        // lambda_x0_b = Q + A^T Lb A - A^T Lb M^{-1} Lb A;   AILM = A^T (I - Lb M^{-1})
        double LbA[X2], LM[X2], ILM[X2], T2[X2], L0[X2], AILM[X2];
        mm<DX, DX, DX>(Lb, A, LbA);
        mm<DX, DX, DX>(Lb, Minv, LM);
#pragma unroll
        for (int i = 0; i < X2; ++i) ILM[i] = (((i / DX) == (i % DX)) ? 1.0 : 0.0) - LM[i];
        mm<DX, DX, DX>(ILM, LbA, T2);  // (I - Lb M^-1) Lb A
#pragma unroll
        for (int i = 0; i < DX; ++i)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = Q[i * DX + j], w = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) {
              s = fma(A[k * DX + i], T2[k * DX + j], s);
              w = fma(A[k * DX + i], ILM[k * DX + j], w);
            }
            L0[i * DX + j] = s;
            AILM[i * DX + j] = w;
          }
        // nu_x0_b = nu_z1_f + AILM (nu_x3_b - Lb a - Lb B mu_u1)
        double Bu[DX], v[DX], n0[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DU; ++k) s = fma(Bm[i * DU + k], mu1[DX + k], s);
          Bu[i] = s;
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = nb[i];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(-Lb[i * DX + k], av[k] + Bu[k], s);
          v[i] = s;
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = nz1[i];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(AILM[i * DX + k], v[k], s);
          n0[i] = s;
        }
        // psi = gamma Lb (sig_x2 (lambda_x2 + inv(inv(Lb) + sig_u2)));  nu_x2_b = lambda_x2_b inv(Lb) nu_x3_b - B mu_u1
        double S3b[X2], Lx2b[X2], T3[X2], T4[X2], gL[X2], psi[X2], nx2b[DX];
#pragma unroll
        for (int i = 0; i < X2; ++i) S3b[i] = Lb[i];
        inv_gj<DX>(S3b);
#pragma unroll
        for (int i = 0; i < X2; ++i) Lx2b[i] = S3b[i] + Su2[i];
        inv_gj<DX>(Lx2b);
#pragma unroll
        for (int i = 0; i < X2; ++i) T3[i] = Lx2[i] + Lx2b[i];
        mm<DX, DX, DX>(Sx2, T3, T4);
        mm<DX, DX, DX>(gamma, Lb, gL);
        mm<DX, DX, DX>(gL, T4, psi);
        mm<DX, DX, DX>(Lx2b, S3b, T3);
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = -Bu[i];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(T3[i * DX + k], nb[k], s);
          nx2b[i] = s;
        }
        // K = -sig_u B^T psi A;  k = sig_u (nu_u_0 + Rug + B^T (gamma nu_x3_b + (I - gamma) nu_x2_b - psi a))
        double w2[DX], Su[DU * DU], BtPsiA[DU * DX], PA[X2];
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DX; ++k) {
            s = fma(gamma[i * DX + k], nb[k], s);
            s = fma((((i == k) ? 1.0 : 0.0) - gamma[i * DX + k]), nx2b[k], s);
            s = fma(-psi[i * DX + k], av[k], s);
          }
          w2[i] = s;
        }
        mm<DX, DX, DX>(psi, A, PA);
#pragma unroll
        for (int r = 0; r < DU; ++r) {
#pragma unroll
          for (int q = 0; q < DU; ++q) Su[r * DU + q] = po[(LY::P_SIG + six(DX + r, DX + q)) * TILE];
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(Bm[k * DU + r], PA[k * DX + j], s);
            BtPsiA[r * DX + j] = s;
          }
        }
#pragma unroll
        for (int r = 0; r < DU; ++r) {
          double kk = 0.0;
#pragma unroll
          for (int q = 0; q < DU; ++q) {
            double inner = nu_u0[q] + Rug[q];
#pragma unroll
            for (int k = 0; k < DX; ++k) inner = fma(Bm[k * DU + q], w2[k], inner);
            kk = fma(Su[r * DU + q], inner, kk);
          }
          po[(LY::P_KK + r) * TILE] = kk;
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < DU; ++q) s = fma(-Su[r * DU + q], BtPsiA[q * DX + j], s);
            po[(LY::P_K + r * DX + j) * TILE] = s;
          }
#pragma unroll
          for (int q = 0; q <= r; ++q) po[(LY::P_SIGK + tix(r, q)) * TILE] = Su[r * DU + q];
        }
#pragma unroll
        for (int i = 0; i < X2; ++i) {
          rr[(LY::R_L0 + i) * TILE] = L0[i];
          Lb[i] = L0[i];
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          rr[(LY::R_N0 + i) * TILE] = n0[i];
          nb[i] = n0[i];
        }
      }
    }
  }

  __device__ __forceinline__ bool load_x0(Carry<DX>& c) {
    const double* q = p.x0 + ((size_t)tile * DX) * TILE + lane;
#pragma unroll
    for (int i = 0; i < DX; ++i) c.m[i] = q[i * TILE];
    q = p.sig_x0 + ((size_t)tile * TRI(DX)) * TILE + lane;
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) {
      c.S[i] = q[i * TILE];
      c.L[i] = c.S[i];
    }
    return chol_rows<DX>(c.L, c.invd);
  }

  __device__ __forceinline__ void metric(int m, int it, double v) const {
    p.metrics[((size_t)m * p.max_iters + it) * p.Bpad + b] = v;
  }

  // compute_update_alpha(update_alpha=True) (i2c.py:921-963): alpha_hat from the summed traces, ratio clip, metrics
  __device__ __forceinline__ double mstep_alpha(int it, double tr, double tr_term, double alpha) {
    double sf = (double)(DZ * p.T);
    if (Env::HAS_TERM && p.has_qf) {
      tr += tr_term;
      sf += (double)DZT;
    }
    double a_new = tr / sf;
    metric(I2C_M_ALPHA_DESIRED, it, a_new);
    if (a_new != a_new) {
      fail(I2C_FAIL_NAN_ALPHA, it, 0);
      a_new = alpha;
    } else if (p.alpha_tol >= 0.0) {
      const double ratio = a_new / alpha;
      const double upper = 2.0 - p.alpha_tol;
      double upd = a_new;
      if (ratio < p.alpha_tol) upd = p.alpha_tol * alpha;
      if (ratio > upper) upd = upper * alpha;
      a_new = upd;
    } else {
      a_new = alpha;
    }
    metric(I2C_M_ALPHA, it, a_new);
    return a_new;
  }

  // ---------------------------------------------------------------------------------- the EM loop
  // TEAM = false: one warp does everything for its tile.  TEAM = true (latency regime, one block of W warps per
  // tile): warp 0 runs the sequential recursions (forward sweep, RTS heads, propagate); the per-cell work of the
  // backward pass that does not feed the recursion (backward_tail) runs on the W-1 helper warps WHILE warp 0 walks the
  // RTS heads (static cell -> warp map, progress counter in shared memory), and the M-step statistics are reduced through
  // shared memory in a fixed order.
  // [it0, it1): the iterations this call runs.  it0 > 0 (em_ticket_kernel: one iteration of one tile per work item) rebuilds
  // the loop-carried state of a launch that started at iteration 0: record roles, cleared independence flags, covariance-control
  // temperature, validity of the per-cell alphas; alpha and the status words travel through global memory.
  template <bool TEAM>
  __device__ void run_impl(const int w, const int W, double* red, const int it0, const int it1) {
    const bool main_warp = !TEAM || w == 0;
    // PROD: the last warp of the team is the copy warp of the record ring (see ring_init); H = warps that run tails
    constexpr bool PROD = TEAM && HOT;
    const bool copy_warp = PROD && w == W - 1;
    const int H = PROD ? W - 2 : W - 1;
    if (main_warp && !PROD && !pipe_ready) pipe_init();
    const double HALF_LOG_2PIE = 1.4189385332046727;  // 0.5 * log(2 pi e)
    double alpha = p.alpha[b];
    const bool aux = !HOT && (p.phases & I2C_PH_STORE_AUX);
    bool flipped = false;  // _update_priors has cleared state_action_independence for index <= tau
    double temp = p.temp0;
    const int T = p.T;
    if (it0 > 0) {
      if (p.phases & I2C_PH_BACKWARD) {
        if (p.phases & I2C_PH_UPDATE_PRIORS) {
          if (it0 & 1) {
            prior = p.post;
            post = p.prior;
          }
          latest = prior;
        } else {
          latest = post;
        }
        if (p.cov_ctrl)
          for (int i = 0; i < it0; ++i) temp += p.dtemp;
      }
      flipped = (p.phases & I2C_PH_UPDATE_PRIORS) != 0;
      if (p.phases & I2C_PH_MSTEP) own_alpha_valid = false;
    }
    for (int it = it0; it < it1; ++it) {
      Carry<DX> c;
      LogAcc ent_x;
      ent_x.reset();
      Stats st;
      st.cost = st.cost_var = st.tr = 0.0;
      st.ent_u.reset();
      double tr_term = 0.0;
      if (main_warp && (p.phases & I2C_PH_FORWARD)) {
        if (!load_x0(c)) fail(I2C_FAIL_CHOL_PRIOR, it, 0);
        if constexpr (PROD) {
          int t = 0;
          if (sweep_is_plain(flipped)) {
            TrigT octx;
            if constexpr (Env::OBS_NL > 0) obs_trig(c, octx);
            Cursor fc = cursor(p.filt, 0, LY::E_FILT);
            for (; t < T - 1; ++t) {
              if constexpr (RCOPY) {
                double pr[LY::E_STAGE_POST];
                ring_read<LY::E_STAGE_POST>(pr);
                forward_cell<true, 1, JOFF>(it, t, 0, alpha, aux, pr, c, ent_x, &octx, fc.ptr);
              } else {
                forward_cell<true, TILE, JOFF>(it, t, 0, alpha, aux, ring_view(), c, ent_x, &octx, fc.ptr);
                ring_release();
              }
              fc.next();
            }
          }
          for (; t < T; ++t) {
            if constexpr (RCOPY) {
              double pr[LY::E_STAGE_POST];
              ring_read<LY::E_STAGE_POST>(pr);
              forward_cell<false, 1, JOFF>(it, t, staged_flags(nullptr, t, flipped), alpha, aux, pr, c, ent_x);
            } else {
              forward_cell<false, TILE, JOFF>(it, t, staged_flags(nullptr, t, flipped), alpha, aux, ring_view(), c, ent_x);
              ring_release();
            }
          }
          // the filtered records written above are read back by the copy warp's bulk copies after the barrier below
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");
        } else {
          stream_begin<LY::E_STAGE_POST>(prior, LY::E_POST, 0);
          for (int t = 0; t < T; ++t) {
            const double* cur = stream<LY::E_STAGE_POST>(prior, LY::E_POST, t, t + 1, t + 1 < T);
            forward_cell(it, t, staged_flags(cur, t, flipped), alpha, aux, cur, c, ent_x);
          }
          stream_end();
        }
      }
      if (copy_warp && (p.phases & I2C_PH_FORWARD)) {
        if (lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          for (int t = 0; t < T; ++t) ring_produce(prior, LY::E_POST, LY::E_STAGE_POST, t);
        }
        __syncwarp();
      }
      if constexpr (PROD && JOFF) {
        if (w == 1 && (p.phases & I2C_PH_FORWARD)) jdyn_serve(T);
      }
      if (p.phases & I2C_PH_BACKWARD) {
        double m3m[DX], S3m[TRI(DX)];
        // TEAM: the tails run CONCURRENTLY with the RTS heads.  Warp 0 publishes the number of finished heads in shared
        // memory (after a block-scope fence: the tails read the posterior from the post records); helper warp w takes the
        // cells i = T-1-t with i mod H = w-1 and waits for head i; the last n_main cells are kept for warp 0, which
        // joins once its heads are done.  The assignment is static, so the per-warp partial sums -- and with them
        // alpha -- are bit-reproducible.  (Before: all heads, a barrier, then all tails: the tails were 11 % of an iteration.)
        const unsigned prog = TEAM ? smem_addr(red + (size_t)7 * W * TILE) : 0u;
        int n_main = 0;
        if constexpr (TEAM) {
          // tail : head cost is about r : 1 (7 with the per-thread record stream, 4 with the copy warp: profiles/r02); balance warp 0's
          // share so that it and the H helpers finish together
          const int r = PROD ? 4 : 7;
          const int nm = (T * (r - H)) / (r * (H + 1));
          n_main = nm > 0 ? nm : 0;
          if (main_warp && lane == 0) progress_publish(prog, 0);
          __syncthreads();  // forward sweep done; progress counter reset
        }
        if (main_warp) {
          if (!(p.phases & I2C_PH_FORWARD)) {
            // resume from the stored filtered message of the last cell
            const double* fr = rec(p.filt, T - 1, LY::E_FILT);
#pragma unroll
            for (int i = 0; i < DX; ++i) c.m[i] = fr[(LY::F_MU3 + i) * TILE];
#pragma unroll
            for (int i = 0; i < TRI(DX); ++i) {
              c.S[i] = fr[(LY::F_SIG3 + i) * TILE];
              c.L[i] = c.S[i];
            }
            chol_rows<DX>(c.L, c.invd);
          }
          backward_terminal(it, T - 1, temp, cell_alpha(T - 1, p.cell_flags[slot(T - 1)], alpha), c, m3m, S3m, tr_term);
          if constexpr (PROD) {
            Cursor pc = cursor(post, T - 1, LY::E_POST);
            for (int t = T - 1; t >= 0; --t) {
              double fr[RCOPY ? LY::E_FILT : 1];
              const double* frv = fr;
              if constexpr (RCOPY) ring_read<LY::E_FILT>(fr);
              else frv = ring_view();
              // publish the PREVIOUS head here: its stores were issued a whole cell ago, so the fence does not wait for them
              // (fencing right after a cell's own stores cost ~90 cycles per cell)
              __syncwarp();
              if (lane == 0) progress_publish(prog, T - 1 - t);
              double mu[N], Sig[TRI(N)];
              backward_head<RCOPY ? 1 : TILE, true>(it, t, aux, frv, m3m, S3m, mu, Sig, pc.ptr);
              pc.prev();
              if constexpr (!RCOPY) ring_release();
            }
            __syncwarp();
            if (lane == 0) progress_publish(prog, T);
          } else if constexpr (TEAM) {
            // RTS heads only: a head is a few hundred cycles, so the filtered records are streamed TEAM_DEPTH cells
            // ahead (a one-cell double buffer would expose the DRAM latency of every record)
            if constexpr (LY::STAGED) {
              constexpr int DEPTH = LY::TEAM_DEPTH;
              static_assert(DEPTH <= kNumBars, "not enough mbarriers");
              rec_fence();
#pragma unroll
              for (int k = 0; k < DEPTH; ++k) {
                const int tt = T - 1 - k;
                if (tt >= 0)
                  rec_issue<LY::E_FILT>(stage + (tt % DEPTH) * (LY::E_FILT * TILE), rec(p.filt, tt, LY::E_FILT), tt % DEPTH);
                if constexpr (!BULK) stage_commit();
              }
              for (int t = T - 1; t >= 0; --t) {
                if constexpr (!BULK) stage_wait<DEPTH - 1>();
                rec_wait(t % DEPTH);
                double* cur = stage + (t % DEPTH) * (LY::E_FILT * TILE);
                double mu[N], Sig[TRI(N)];
                backward_head(it, t, aux, cur, m3m, S3m, mu, Sig);
                if (t - DEPTH >= 0) rec_issue<LY::E_FILT>(cur, rec(p.filt, t - DEPTH, LY::E_FILT), t % DEPTH);
                if constexpr (!BULK) stage_commit();
                __syncwarp();
                if (lane == 0) progress_publish(prog, T - t);
              }
              if constexpr (!BULK) stage_wait<0>();
            } else {
              for (int t = T - 1; t >= 0; --t) {
                double mu[N], Sig[TRI(N)];
                backward_head(it, t, aux, rec(p.filt, t, LY::E_FILT), m3m, S3m, mu, Sig);
                __syncwarp();
                if (lane == 0) progress_publish(prog, T - t);
              }
            }
          } else {
            stream_begin<LY::E_FILT>(p.filt, LY::E_FILT, T - 1);
            for (int t = T - 1; t >= 0; --t) {
              const double* cur = stream<LY::E_FILT>(p.filt, LY::E_FILT, t, t - 1, t > 0);
              backward_cell(it, t, staged_flags(cur, t, false), aux, cur, m3m, S3m, st);
            }
            stream_end();
          }
        }
        if (copy_warp) {
          if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            for (int t = T - 1; t >= 0; --t) ring_produce(p.filt, LY::E_FILT, LY::E_FILT, t);
          }
          __syncwarp();
        }
        if (p.cov_ctrl) temp += p.dtemp;
        if constexpr (TEAM) {
          // cells [0, T - n_main) (in head order i = T-1-t) belong to the helpers, the rest to warp 0
          const int i0 = main_warp ? T - n_main : w - 1, di = main_warp ? 1 : H;
          const int i1 = main_warp ? T : (copy_warp ? 0 : T - n_main);
          if (W == 4 && N <= 3) {  // (n = 5: the 20 prefetch registers make the tail spill -- cart-pole 8192 x 200: 0.86 -> 0.98 ms)
            // 4-warp teams (two blocks per SM) are bound by their two tail warps, and ~40 % of a tail was the L2 round trip
            // of the posterior it starts with (profiles/r02l source view: long_scoreboard).  Software pipeline: the posterior
            // of this warp's NEXT cell is requested before the current tail is computed -- whenever its head has been
            // published already, which is the rule when the tails are the bottleneck.
            double mun[N], Sign[TRI(N)];
            bool have = false;
            for (int i = i0; i < i1; i += di) {
              const int t = T - 1 - i;
              double mu[N], Sig[TRI(N)];
              if (have) {
#pragma unroll
                for (int k = 0; k < N; ++k) mu[k] = mun[k];
#pragma unroll
                for (int k = 0; k < TRI(N); ++k) Sig[k] = Sign[k];
              } else {
                if (!main_warp) {
                  while (progress_read(prog) <= i) __nanosleep(64);
                }
                const double* po = rec(post, t, LY::E_POST);
#pragma unroll
                for (int k = 0; k < N; ++k) mu[k] = __ldcg(po + (LY::P_MU + k) * TILE);
#pragma unroll
                for (int k = 0; k < TRI(N); ++k) Sig[k] = __ldcg(po + (LY::P_SIG + k) * TILE);
              }
              const int in = i + di;
              have = in < i1 && (main_warp || progress_read(prog) > in);
              if (have) {
                const double* pn = rec(post, T - 1 - in, LY::E_POST);
#pragma unroll
                for (int k = 0; k < N; ++k) mun[k] = __ldcg(pn + (LY::P_MU + k) * TILE);
#pragma unroll
                for (int k = 0; k < TRI(N); ++k) Sign[k] = __ldcg(pn + (LY::P_SIG + k) * TILE);
              }
              backward_tail(it, t, aux, nullptr, mu, Sig, st);
            }
          } else {
          for (int i = i0; i < i1; i += di) {
            if (!main_warp) {
              while (progress_read(prog) <= i) __nanosleep(64);
            }
            const int t = T - 1 - i;
            const double* po = rec(post, t, LY::E_POST);
            double mu[N], Sig[TRI(N)];
#pragma unroll
            for (int k = 0; k < N; ++k) mu[k] = __ldcg(po + (LY::P_MU + k) * TILE);
#pragma unroll
            for (int k = 0; k < TRI(N); ++k) Sig[k] = __ldcg(po + (LY::P_SIG + k) * TILE);
            backward_tail(it, t, aux, nullptr, mu, Sig, st);
          }
          }
          if constexpr (BULK) {  // K, k, sigK written here are read back by warp 0's bulk copies in the next forward sweep
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          // fixed-order reduction of the per-warp statistics: red[k][w][lane]
          double* r = red + (size_t)w * TILE + lane;
          r[0 * W * TILE] = st.cost;
          r[1 * W * TILE] = st.cost_var;
          r[2 * W * TILE] = st.tr;
          r[3 * W * TILE] = st.ent_u.m;
          r[4 * W * TILE] = (double)st.ent_u.e;
          r[5 * W * TILE] = (double)status;
          r[6 * W * TILE] = (double)info;
          __syncthreads();
          if (main_warp) {
            st.cost = st.cost_var = st.tr = 0.0;
            st.ent_u.reset();
            for (int ww = 0; ww < W; ++ww) {
              const double* q = red + (size_t)ww * TILE + lane;
              st.cost += q[0 * W * TILE];
              st.cost_var += q[1 * W * TILE];
              st.tr += q[2 * W * TILE];
              st.ent_u.mul(q[3 * W * TILE]);
              st.ent_u.e += (int)q[4 * W * TILE];
              if (status == I2C_OK && q[5 * W * TILE] != 0.0) {
                status = (int)q[5 * W * TILE];
                info = (int)q[6 * W * TILE];
              }
            }
          }
        }
        latest = post;
      }
      PStats ps;
      ps.cost = ps.cost_var = ps.tr = 0.0;
      ps.cost_min = INFINITY;
      ps.ent.reset();
      if (main_warp && (p.phases & I2C_PH_PROPAGATE)) {
        Carry<DX> cp;
        if (!load_x0(cp)) fail(I2C_FAIL_CHOL_PROPAGATE, it, 0);
        if constexpr (PROD) {
          for (int t = 0; t < T; ++t) {
            if constexpr (RCOPY) {
              double po[LY::E_STAGE_POST];
              ring_read<LY::E_STAGE_POST>(po);
              propagate_cell<1>(it, t, staged_flags(nullptr, t, flipped), aux, po, cp, ps);
            } else {
              propagate_cell(it, t, staged_flags(nullptr, t, flipped), aux, ring_view(), cp, ps);
              ring_release();
            }
          }
        } else {
          stream_begin<LY::E_STAGE_POST>(latest, LY::E_POST, 0);
          for (int t = 0; t < T; ++t) {
            const double* cur = stream<LY::E_STAGE_POST>(latest, LY::E_POST, t, t + 1, t + 1 < T);
            propagate_cell(it, t, staged_flags(cur, t, flipped), aux, cur, cp, ps);
          }
          stream_end();
        }
        if (p.cov_ctrl) {
          // KL(N(mu_x3_pf, sig_x3_pf) || N(mu_xT, sig_xT)) of the last cell (i2c.py:1012-1019, 1223-1229)
          double A[TRI(DX)], ia[DX], d[DX];
#pragma unroll
          for (int i = 0; i < TRI(DX); ++i) A[i] = p.sxt[i];
          chol_rows<DX>(A, ia);
#pragma unroll
          for (int i = 0; i < DX; ++i) d[i] = p.mu_xt[i] - cp.m[i];
          fwd_subst<DX>(A, ia, d);
          double dist = 0.0, tr = 0.0, ld1 = 0.0;
#pragma unroll
          for (int i = 0; i < DX; ++i) dist = fma(d[i], d[i], dist);
          // tr(SigT^{-1} Sig1) = || La^{-1} L1 ||_F^2
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double w[DX];
#pragma unroll
            for (int i = 0; i < DX; ++i) w[i] = (i >= j) ? cp.L[tix(i, j)] : 0.0;
            fwd_subst<DX>(A, ia, w);
#pragma unroll
            for (int i = 0; i < DX; ++i) tr = fma(w[i], w[i], tr);
            ld1 += log(cp.L[tix(j, j)]);
          }
          metric(I2C_M_KL_TERM, it, 0.5 * (p.sxt_logdet - 2.0 * ld1 + tr + dist - (double)DX));
        }
      }
      if (copy_warp && (p.phases & I2C_PH_PROPAGATE)) {
        if (lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          for (int t = 0; t < T; ++t) ring_produce(latest, LY::E_POST, LY::E_STAGE_POST, t);
        }
        __syncwarp();
      }
      if (main_warp && (p.phases & I2C_PH_RICCATI)) riccati_sweep(alpha);
      if (main_warp && (p.phases & I2C_PH_MSTEP)) {
        metric(I2C_M_COST_M, it, st.cost);
        metric(I2C_M_COST_M_VAR, it, st.cost_var);
        if (p.phases & I2C_PH_PROPAGATE) {
          metric(I2C_M_COST_PF, it, ps.cost);
          metric(I2C_M_COST_PF_VAR, it, ps.cost_var);
          metric(I2C_M_COST_PF_MIN, it, ps.cost_min);
          metric(I2C_M_ALPHA_PF, it, ps.tr / (double)(DZ * T));
          metric(I2C_M_PROPAGATE_ENTROPY, it, (double)(T * DX) * HALF_LOG_2PIE + ps.ent.value());
        } else {
          metric(I2C_M_COST_PF, it, -1.0);
        }
        metric(I2C_M_POLICY_ENTROPY, it, (double)(T * DU) * HALF_LOG_2PIE + st.ent_u.value());
        metric(I2C_M_X_PRIOR_ENTROPY, it, (double)(T * DX) * HALF_LOG_2PIE + ent_x.value());
      }
      if (p.phases & I2C_PH_UPDATE_PRIORS) {
        // _update_priors (i2c.py:1210-1221): prior <- posterior (record swap), independence cleared for index <= tau
        if (latest == post) {
          double* tmp = prior;
          prior = post;
          post = tmp;
        }
        flipped = true;
      }
      if (main_warp && (p.phases & I2C_PH_MSTEP)) {
        alpha = mstep_alpha(it, st.tr, tr_term, alpha);
        own_alpha_valid = false;  // update_xi pushes the new sig_xi to every cell (i2c.py:976-981)
      }
      if (main_warp && (p.phases & I2C_PH_CALIBRATE)) {
        // calibrate_alpha (i2c.py:895-911): alpha from the propagated cost features, no terminal term
        double a_pf = ps.tr / (double)(DZ * T);
        bool upd = (p.phases & I2C_PH_ONLY_DECREASE) ? (a_pf < alpha) : true;
        if (upd) {
          alpha = a_pf;
          own_alpha_valid = false;
        }
        metric(I2C_M_ALPHA, it, alpha);
      }
    }
    if (!main_warp) return;
    p.alpha[b] = alpha;
    if (status != I2C_OK && p.status[b] == I2C_OK) {
      p.status[b] = status;
      p.info[b] = info;
    }
  }
  __device__ void run() { run_impl<false>(0, 1, nullptr, 0, p.n_iter); }
  __device__ void run_one(int it) { run_impl<false>(0, 1, nullptr, it, it + 1); }
};

// Opt a kernel in to > 48 KB of dynamic shared memory.  The attribute is per DEVICE: remembered per device ordinal (a
// process-wide flag made the launch fail on every device after the first one).
template <auto Kernel>
static int allow_big_smem() {
  static unsigned long long done = 0;  // one bit per device ordinal
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 64 && ((done >> dev) & 1ull)) return 0;
  e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  if (dev < 64) done |= 1ull << dev;
  return 0;
}

// MINB = minimum resident 128-thread blocks per SM: 1 lets ptxas use up to 255 registers (best per-warp latency,
// used when the batch cannot fill the machine anyway); 4 caps the kernel at 128 registers so that 16 warps per SM
// are resident (throughput regime, small envs only).
template <class Env, int MINB, bool LAT, bool LIN, bool GH>
__global__ void __launch_bounds__(128, MINB) em_kernel(const __grid_constant__ KParams pin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / TILE;
  const int lane = threadIdx.x % TILE;
  if (warp >= pin.ntiles) return;
  extern __shared__ __align__(128) double stage_smem[];
  constexpr int PER_WARP = 2 * Lay<Env>::E_STAGE_TOT * TILE + kNumBars;  // staging buffers + mbarriers
  double* base = stage_smem + (size_t)(threadIdx.x / TILE) * PER_WARP;
  Worker<Env, LAT, LIN, GH> w(pin, warp, lane, base + lane, reinterpret_cast<uint64_t*>(base + 2 * Lay<Env>::E_STAGE_TOT * TILE));
  w.run();
}

// Throughput regime without wave quantisation: a work item is ONE EM iteration of one tile; the resident warps (148 x 3
// blocks x 4 warps = one full wave of the spill-free 168-register variant) draw items from a ticket counter, ticket k =
// (iteration k / ntiles, tile k % ntiles).  Iteration `it` of a tile needs iteration it-1 of the same tile: its ticket was
// drawn ntiles tickets earlier, i.e. more than a full wave ago, so the acquire-spin on the tile's finished-iteration counter
// almost never waits; it cannot deadlock (tickets are only held by running warps and only wait for smaller tickets).
// A batch of 65 536 problems x 20 iterations is 40 960 items on 1 776 warps = 23.06 rounds instead of two waves of the
// whole 20-iteration job at 115 % / 15 % fill.  The loop-carried state is rebuilt per item (Worker::run_impl, it0 > 0); records
// written by the previous item of the tile -- possibly on another SM -- are ordered by fence + release / acquire at gpu scope.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <class Env>
__global__ void __launch_bounds__(128, 3) em_ticket_kernel(const __grid_constant__ KParams pin) {
  const int lane = threadIdx.x % TILE;
  extern __shared__ __align__(128) double stage_smem[];
  constexpr int PER_WARP = 2 * Lay<Env>::E_STAGE_TOT * TILE + kNumBars;  // staging buffers + mbarriers
  double* base = stage_smem + (size_t)(threadIdx.x / TILE) * PER_WARP;
  int* const ticket = pin.tickets;
  int* const done = pin.tickets + 1;
  const int total = pin.ntiles * pin.n_iter;
  unsigned phase = 0;
  bool inited = false;
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= total) break;
    const int it = k / pin.ntiles, tile = k - it * pin.ntiles;
    if (it > 0) {
      if (lane == 0)
        while (ld_acquire_gpu(done + tile) < it) __nanosleep(256);
      __syncwarp();
      asm volatile("fence.proxy.async;" ::: "memory");  // the records are read back with bulk copies
    }
    {
      // the warp's mbarriers are initialised by its first item and keep their phases across items (re-initialising them per
      // item is legal -- every copy has been waited for -- but hides the ordering from compute-sanitizer's racecheck)
      Worker<Env, false> w(pin, tile, lane, base + lane, reinterpret_cast<uint64_t*>(base + 2 * Lay<Env>::E_STAGE_TOT * TILE));
      w.bar_phase = phase;
      w.pipe_ready = inited;
      w.run_one(it);
      phase = w.bar_phase;
      inited = true;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_gpu(done + tile, it + 1);
  }
}

// Latency-regime kernel: one block of W warps per tile (see Worker::run_impl<true>).
template <class Env, int W, int HOT>
__global__ void __launch_bounds__(W * TILE, 8 / W) em_team_kernel(const __grid_constant__ KParams pin) {
  // (W = 4, two blocks per SM: warp k of either block sits on sub-partition k, so both main warps share sub-partition 0.
  // Tried in round 2: rotating the second block's warp roles by two (its main warp on sub-partition 2, next to a tail warp of
  // the first block): 0.337 instead of 0.328 ms per iteration at 8192 pendulum problems; reversing them (each main warp next to
  // the other block's copy warp, whose mbarrier polling takes issue slots): 0.411 ms.  The W = 4 regime is bound by its two
  // tail warps per tile, not by the main warps' placement.)
  const int w = threadIdx.x / TILE, lane = threadIdx.x % TILE;
  extern __shared__ __align__(128) double stage_smem[];
  double* red = stage_smem + Lay<Env>::E_TEAM_STAGE * TILE + kNumBars;
  Worker<Env, true, false, false, HOT> wk(pin, blockIdx.x, lane, stage_smem + lane,
                       reinterpret_cast<uint64_t*>(stage_smem + Lay<Env>::E_TEAM_STAGE * TILE));
  if constexpr (HOT) {
    if (threadIdx.x == 0) wk.ring_init();
    // cell targets and {flags, index} of the whole horizon, in logical cell order (the ring offset is fixed during a launch)
    constexpr int DZ = Env::DZ;
    double* ztab = red + (size_t)7 * W * TILE + 2;
    int2* ftab = reinterpret_cast<int2*>(ztab + (size_t)pin.T * DZ);
    for (int i = threadIdx.x; i < pin.T * DZ; i += W * TILE) ztab[i] = pin.z_cell[wk.slot(i / DZ) * DZ + i % DZ];
    for (int t = threadIdx.x; t < pin.T; t += W * TILE) ftab[t] = make_int2(pin.cell_flags[wk.slot(t)], pin.cell_index[wk.slot(t)]);
    __syncthreads();
    wk.ztab_s = smem_addr(ztab);
    wk.ftab_s = smem_addr(ftab);
    using WK = Worker<Env, true, false, false, HOT>;
    if constexpr (WK::JOFF) {
      // [2][JELEMS][32] hand-over slots in the tail of the staging area (the ring takes RING * E_STAGE of its E_TEAM_STAGE rows)
      static_assert(Lay<Env>::RING * Lay<Env>::E_STAGE + 2 * WK::JELEMS <= Lay<Env>::E_TEAM_STAGE, "no room for the J_dyn slots");
      wk.jslot_s = smem_addr(stage_smem + (size_t)Lay<Env>::RING * Lay<Env>::E_STAGE * TILE + lane);
    }
  }
  wk.template run_impl<true>(w, W, red, 0, pin.n_iter);
}

template <class Env, int W, int HOT>
static int launch_em_team_v(const KParams& p, cudaStream_t s) {
  // staging + mbarriers + reduction + progress counter (+ HOT: target / flag table of the horizon)
  const size_t smem = ((size_t)(Lay<Env>::E_TEAM_STAGE + 7 * W) * TILE + kNumBars + 2 + (HOT ? (size_t)p.T * (Env::DZ + 1) : 0)) * sizeof(double);

  if (int e = allow_big_smem<em_team_kernel<Env, W, HOT>>()) return e;
  KParams q = p;
  q.stage_meta = Lay<Env>::STAGED;
  em_team_kernel<Env, W, HOT><<<p.ntiles, W * TILE, smem, s>>>(q);
  return (int)cudaGetLastError();
}
template <class Env, int W>
static int launch_em_team(const KParams& p, cudaStream_t s) {
  if constexpr (Env::DY > 0) {
    if (p.hot == 2) return launch_em_team_v<Env, W, 2>(p, s);
  }
  if (p.hot == 1) return launch_em_team_v<Env, W, 1>(p, s);
  return launch_em_team_v<Env, W, 0>(p, s);
}

template <class Env, int MINB, bool LAT, bool LIN = false, bool GH = false>
static int launch_em_v(const KParams& p, cudaStream_t s, int threads) {
  int wpb = threads / TILE;
  int blocks = (p.ntiles + wpb - 1) / wpb;
  const size_t smem = Lay<Env>::STAGED ? (size_t)wpb * (2 * Lay<Env>::E_STAGE_TOT * TILE + kNumBars) * sizeof(double) : 0;
  if (int e = allow_big_smem<em_kernel<Env, MINB, LAT, LIN, GH>>()) return e;
  KParams q = p;
  // staging the per-cell targets / flags removes their exposed load latency when a warp is alone on its
  // sub-partition; in the throughput regime the extra LDGSTS instructions cost more than they hide
  q.stage_meta = Lay<Env>::STAGED && LAT;
  em_kernel<Env, MINB, LAT, LIN, GH><<<blocks, threads, smem, s>>>(q);
  return (int)cudaGetLastError();
}

template <class Env>
static int launch_em_ticket(const KParams& p, cudaStream_t s) {
  constexpr int WPB = 4, RESIDENT = 148 * 3;  // blocks of the full wave
  const size_t smem = Lay<Env>::STAGED ? (size_t)WPB * (2 * Lay<Env>::E_STAGE_TOT * TILE + kNumBars) * sizeof(double) : 0;
  if (int e = allow_big_smem<em_ticket_kernel<Env>>()) return e;
  cudaError_t ce = cudaMemsetAsync(p.tickets, 0, (size_t)(p.ntiles + 1) * sizeof(int32_t), s);
  if (ce != cudaSuccess) return (int)ce;
  KParams q = p;
  q.stage_meta = 0;
  const long long items = (long long)p.ntiles * p.n_iter;
  const int blocks = (int)((items + WPB - 1) / WPB < RESIDENT ? (items + WPB - 1) / WPB : RESIDENT);
  em_ticket_kernel<Env><<<blocks, WPB * TILE, smem, s>>>(q);
  return (int)cudaGetLastError();
}

constexpr int kGroupNotTaken = -12345;
template <class Env>
static int launch_em_group_maybe(const KParams& p, cudaStream_t s);  // i2c_group.cuh

template <class Env>
static int launch_em_t(const KParams& p, cudaStream_t s) {
  // one warp per tile of 32 problems; small blocks spread the warps over all SMs / sub-partitions
  int threads = 32;
  if (p.ntiles >= 148 * 8) threads = 64;
  if (p.ntiles >= 148 * 32) threads = 128;
  // Linearize inference: one (latency-style) variant only -- not a throughput path
  if (p.linearize) return launch_em_v<Env, 1, true, true>(p, s, threads);
  // Gauss-Hermite grids: degree^n points per transform, one variant as well
  if (p.gh.degree > 0) return launch_em_v<Env, 1, true, false, true>(p, s, threads);
  // A/B and tests: I2C_B200_MINB = 3 / 4 / 5 forces a throughput variant (5 = ticket kernel) whatever the batch size
  if constexpr (Lay<Env>::N <= 3) {
    if (p.minb == 5 && p.tickets && !(p.phases & I2C_PH_CALIBRATE)) return launch_em_ticket<Env>(p, s);
    if (p.minb == 3) return launch_em_v<Env, 3, false>(p, s, 128);
    if (p.minb == 4) return launch_em_v<Env, 4, false>(p, s, 128);
  }
  // small batches: G lanes per problem (i2c_group.cuh); the launcher there decides
  {
    const int rc = launch_em_group_maybe<Env>(p, s);
    if (rc != kGroupNotTaken) return rc;
  }
  // fewer tiles than SMs: spread each tile's backward pass over the 8 warps of a block (one block per SM)
  if (p.ntiles <= 148 && p.T >= 8 && !p.no_team) return launch_em_team<Env, 8>(p, s);
  if (p.ntiles <= 296 && p.T >= 8 && !p.no_team) return launch_em_team<Env, 4>(p, s);  // two blocks per SM
  if constexpr (Lay<Env>::N <= 3) {
    // small envs fit 128 registers with a few bytes of spill: worth it once >= 12 warps per SM are available
    if (p.ntiles >= 148 * 12) {
      // 168 registers / 12 warps per SM (no spills) is 17 % faster per wave than 128 registers / 16 warps per SM, but its
      // waves hold 1776 warps instead of 2368: take it when the batch fills its waves to >= 87 % (measured: 10.3e9 against
      // 8.8e9 updates/s at 56 832 and 113 664 problems, 10.3e9 against 9.7e9 at 262 144; 7.0e9 against 8.9e9 at 65 536)
      const int waves3 = (p.ntiles + 148 * 12 - 1) / (148 * 12);
      const bool fills3 = (long long)waves3 * 148 * 12 * 100 <= (long long)p.ntiles * 115;
      // several iterations in one launch: (tile, iteration) work items from a ticket counter keep the 1776 resident warps of the
      // spill-free variant busy whatever the batch size (em_ticket_kernel).  Estimated rounds, in units of one full 168-register
      // wave: items / 1776 (rounded up) against the static choice below (a 128-register wave holds 2368 warps and takes 1.56 x
      // as long).  (I2C_B200_MINB = 5 forces it, 3 / 4 force the static variants: see above.)
      if (p.tickets && !(p.phases & I2C_PH_CALIBRATE) && p.n_iter >= 2) {
        const long long items = (long long)p.ntiles * p.n_iter;
        const double rounds_ticket = (double)((items + 148 * 12 - 1) / (148 * 12));
        const double rounds_static = fills3 ? (double)waves3 * p.n_iter : 1.56 * ((p.ntiles + 148 * 16 - 1) / (148 * 16)) * p.n_iter;
        if (rounds_ticket <= rounds_static) return launch_em_ticket<Env>(p, s);  // (equal: measured 6 % faster, no wave edges)
      }
      const int minb = fills3 ? 3 : 4;
      if (minb == 3) return launch_em_v<Env, 3, false>(p, s, 128);
      return launch_em_v<Env, 4, false>(p, s, 128);
    }
  }
  // fewer than 4 warps per SM: every warp is latency-bound -> cp.async + staged targets; else TMA bulk copies
  if (p.ntiles < 148 * 4) return launch_em_v<Env, 1, true>(p, s, threads);
  return launch_em_v<Env, 1, false>(p, s, threads);
}

// ------------------------------------------------------------------------------------------------
// Stand-alone sigma-point transform: QuadratureInference.forward / forward_gaussian
// (inference/quadrature.py:27-58) for the registered env maps.
template <class Env, int FN>
struct QuadDims {
  static constexpr int N = Env::DX + Env::DU;
  static constexpr int D = (FN == 0 || FN == 2) ? N : Env::DX;
  static constexpr int DY = FN == 0 ? Env::DZ : (FN == 1 ? Env::DZT : (FN == 2 ? Env::DX : (Env::DY > 0 ? Env::DY : 1)));
};

template <class Env, int FN>
__global__ void __launch_bounds__(64) quad_kernel(const __grid_constant__ QuadArgs a) {
  using QD = QuadDims<Env, FN>;
  constexpr int D = QD::D, DY = QD::DY;
  const int tile = (blockIdx.x * blockDim.x + threadIdx.x) / TILE, lane = threadIdx.x % TILE;
  if (tile >= a.ntiles) return;
  double par[Env::NP > 0 ? Env::NP : 1];
#pragma unroll
  for (int i = 0; i < Env::NP; ++i) par[i] = a.envpar[((size_t)tile * Env::NP + i) * TILE + lane];
  double m[D], L[TRI(D)], invd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) m[i] = a.m[((size_t)tile * D + i) * TILE + lane];
#pragma unroll
  for (int i = 0; i < TRI(D); ++i) L[i] = a.S[((size_t)tile * TRI(D) + i) * TILE + lane];
  bool ok = chol_rows<D>(L, invd);
  double my[DY], Sy[TRI(DY)], Sxy[D * DY];
  typename Env::TrigT ctx;
  Env::center(m, ctx);
  auto ev = [&](const double* x, int j, double* y) {
    if (FN == 0) Env::obs(x, j, ctx, y);
    else if (FN == 1) Env::obs_term(x, j, ctx, y);
    else if (FN == 2) Env::dyn(x, j, ctx, par, y);
    else Env::measure(x, j, ctx, y);
  };
  if (a.gh.degree > 0) grid_transform<D, DY>(m, L, a.gh, ev, my, Sy, Sxy);
  else sigma_transform<D, DY>(m, L, a.sf, a.w0, a.wi, ev, my, Sy, Sxy);
  cross_cov<D, DY>(L, Sxy);
#pragma unroll
  for (int i = 0; i < DY; ++i) a.my[((size_t)tile * DY + i) * TILE + lane] = my[i];
#pragma unroll
  for (int i = 0; i < TRI(DY); ++i) a.Sy[((size_t)tile * TRI(DY) + i) * TILE + lane] = Sy[i];
#pragma unroll
  for (int i = 0; i < D * DY; ++i) a.Sxy[((size_t)tile * D * DY + i) * TILE + lane] = Sxy[i];
  if (a.status) a.status[tile * TILE + lane] = ok ? I2C_OK : I2C_FAIL_CHOL_PRIOR;
}

template <class Env>
static int launch_quad_t(int fn, const QuadArgs& a, cudaStream_t s) {
  int threads = 64, blocks = (a.ntiles * TILE + threads - 1) / threads;
  switch (fn) {
    case 0: quad_kernel<Env, 0><<<blocks, threads, 0, s>>>(a); break;
    case 1: quad_kernel<Env, 1><<<blocks, threads, 0, s>>>(a); break;
    case 2: quad_kernel<Env, 2><<<blocks, threads, 0, s>>>(a); break;
    case 3:
      if (Env::DY == 0) return -2;
      quad_kernel<Env, 3><<<blocks, threads, 0, s>>>(a);
      break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Cubature Kalman filter step: PartiallyObservedMpcPolicy.filter (policy/mpc.py:125-145).
template <class Env>
__global__ void __launch_bounds__(64) ckf_kernel(const __grid_constant__ CkfArgs a) {
  constexpr int DX = Env::DX, DU = Env::DU, N = DX + DU, DY = Env::DY > 0 ? Env::DY : 1;
  const int tile = (blockIdx.x * blockDim.x + threadIdx.x) / TILE, lane = threadIdx.x % TILE;
  if (tile >= a.ntiles) return;
  double par[Env::NP > 0 ? Env::NP : 1];
#pragma unroll
  for (int i = 0; i < Env::NP; ++i) par[i] = a.envpar[((size_t)tile * Env::NP + i) * TILE + lane];
  double m[DX], S[TRI(DX)], L[TRI(DX)], invd[DX], u[DU];
#pragma unroll
  for (int i = 0; i < DX; ++i) m[i] = a.x0[((size_t)tile * DX + i) * TILE + lane];
#pragma unroll
  for (int i = 0; i < TRI(DX); ++i) L[i] = a.sig_x0[((size_t)tile * TRI(DX) + i) * TILE + lane];
  const int bc = min(tile * TILE + lane, a.B - 1);  // canonical inputs: padded lanes repeat the last roll-out
#pragma unroll
  for (int i = 0; i < DU; ++i) u[i] = a.canonical ? a.u[(size_t)bc * DU + i] : a.u[((size_t)tile * DU + i) * TILE + lane];
  bool ok = chol_rows<DX>(L, invd);
  typename Env::TrigT ctx;
  // predict: pass the belief through the dynamics with the applied control appended to every point
  double mf[DX], Dm0[DX * DX];
  Env::center(m, ctx);
  sigma_transform<DX, DX>(m, L, a.sf, a.w0, a.wi,
                          [&](const double* x, int j, double* y) {
                            double xu[N];
#pragma unroll
                            for (int i = 0; i < DX; ++i) xu[i] = x[i];
#pragma unroll
                            for (int i = 0; i < DU; ++i) xu[DX + i] = u[i];
                            Env::dyn(xu, j, ctx, par, y);
                          },
                          mf, S, Dm0);
#pragma unroll
  for (int i = 0; i < TRI(DX); ++i) {
    S[i] += a.sig_eta[i];
    L[i] = S[i];
  }
  ok = chol_rows<DX>(L, invd) && ok;
  // update on the measurement
  double my[DY], Sy[TRI(DY)], Sxy[DX * DY], yv[DY];
  Env::center(mf, ctx);
  sigma_transform<DX, DY>(mf, L, a.sf, a.w0, a.wi,
                          [&](const double* x, int j, double* y) { Env::measure(x, j, ctx, y); }, my, Sy, Sxy);
  cross_cov<DX, DY>(L, Sxy);
#pragma unroll
  for (int i = 0; i < TRI(DY); ++i) Sy[i] += a.sig_zeta[i];
#pragma unroll
  for (int i = 0; i < DY; ++i) yv[i] = a.canonical ? a.y[(size_t)bc * DY + i] : a.y[((size_t)tile * DY + i) * TILE + lane];
  ok = condition<DX, DY>(mf, S, Sy, Sxy, my, yv) && ok;
#pragma unroll
  for (int i = 0; i < DX; ++i) a.x0[((size_t)tile * DX + i) * TILE + lane] = mf[i];
#pragma unroll
  for (int i = 0; i < TRI(DX); ++i) a.sig_x0[((size_t)tile * TRI(DX) + i) * TILE + lane] = S[i];
  if (!ok && a.status[tile * TILE + lane] == I2C_OK) a.status[tile * TILE + lane] = I2C_FAIL_CKF;
}

template <class Env>
static int launch_ckf_t(const CkfArgs& a, cudaStream_t s) {
  if constexpr (Env::DY > 0) {
    int threads = 64, blocks = (a.ntiles * TILE + threads - 1) / threads;
    ckf_kernel<Env><<<blocks, threads, 0, s>>>(a);
    return (int)cudaGetLastError();
  } else {
    return -2;  // only the quadrotor defines measure() (mpc_quad.py:371-383)
  }
}

// ------------------------------------------------------------------------------------------------
// Batched stochastic closed-loop evaluation of time-indexed linear-Gaussian controllers:
// BaseSim.run / batch_eval (i2c/env.py:40-103) with BaseKnownSim.forward (:180-187) under
// TimeIndexedLinearGaussianPolicy / ExpertTimeIndexedLinearGaussianPolicy (i2c/policy/linear.py:31-43, 73-90).
// One thread per (problem, roll-out); SURVEY.md 8(f) row 1.
template <class Env>
__global__ void __launch_bounds__(128) rollout_kernel(const __grid_constant__ RolloutArgs a) {
  constexpr int DX = Env::DX, DU = Env::DU, N = DX + DU, DZ = Env::DZ, DZT = Env::DZT;
  const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (id >= (long long)a.B * a.R) return;
  const int b = (int)(id / a.R);
  double par[Env::NP > 0 ? Env::NP : 1];
#pragma unroll
  for (int i = 0; i < Env::NP; ++i) par[i] = a.envpar[(size_t)b * Env::NP + i];
  curandStatePhilox4_32_10_t rng;
  const bool need_rng = (!a.eta && !a.noise_free) || (a.sigK && !a.eps_u);
  if (need_rng) curand_init(a.seed, (unsigned long long)id, 0, &rng);
  typename Env::TrigT ctx;
  double xu[N];
#pragma unroll
  for (int i = 0; i < DX; ++i) xu[i] = a.x_init[(size_t)id * DX + i];
  for (int t = 0; t < a.T; ++t) {
    const size_t bt = (size_t)b * a.T + t;
    const double* K = a.K + bt * DU * DX;
    double gate = 1.0, d[DX];
    if (a.ex_mu) {  // expert policy: u = k + p K (x - mu), p = exp(-e) or [|e| < threshold], e = 1/2 d^T lam d
      double e = 0.0;
#pragma unroll
      for (int i = 0; i < DX; ++i) d[i] = xu[i] - a.ex_mu[bt * DX + i];
#pragma unroll
      for (int i = 0; i < DX; ++i)
#pragma unroll
        for (int j = 0; j < DX; ++j) e = fma(d[i] * a.ex_lam[(bt * DX + i) * DX + j], d[j], e);
      e *= 0.5;
      gate = a.soft_expert ? exp(-e) : (fabs(e) < a.hard_threshold ? 1.0 : 0.0);
    } else {
#pragma unroll
      for (int i = 0; i < DX; ++i) d[i] = xu[i];
    }
    double u[DU];
#pragma unroll
    for (int r = 0; r < DU; ++r) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < DX; ++i) s = fma(K[r * DX + i], d[i], s);
      u[r] = fma(gate, s, a.k[bt * DU + r]);
    }
    if (a.sigK) {  // sample the action: u += chol(sig_k) eps
      double Ls[TRI(DU)], iv[DU], eps[DU];
#pragma unroll
      for (int r = 0; r < DU; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) Ls[tix(r, q)] = a.sigK[(bt * DU + r) * DU + q];
      const bool pd = chol_rows<DU>(Ls, iv);
#pragma unroll
      for (int r = 0; r < DU; ++r) eps[r] = a.eps_u ? a.eps_u[((size_t)id * a.T + t) * DU + r] : curand_normal_double(&rng);
      if (pd) {
#pragma unroll
        for (int r = 0; r < DU; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) u[r] = fma(Ls[tix(r, q)], eps[q], u[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < DU; ++r) xu[DX + r] = u[r];
    double* oxu = a.xu + ((size_t)id * a.T + t) * N;
#pragma unroll
    for (int i = 0; i < N; ++i) oxu[i] = xu[i];
    double zz[DZ];
    Env::obs(xu, 0, ctx, zz);
#pragma unroll
    for (int i = 0; i < DZ; ++i) a.z[((size_t)id * a.T + t) * DZ + i] = zz[i];
    double xn[DX];
    Env::dyn(xu, 0, ctx, par, xn);
    if (a.eta) {
#pragma unroll
      for (int i = 0; i < DX; ++i) xn[i] += a.eta[((size_t)id * a.T + t) * DX + i];
    } else if (!a.noise_free) {
      double eps[DX];
#pragma unroll
      for (int i = 0; i < DX; ++i) eps[i] = curand_normal_double(&rng);
#pragma unroll
      for (int i = 0; i < DX; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) xn[i] = fma(a.chol_eta[tix(i, j)], eps[j], xn[i]);
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) xu[i] = xn[i];
  }
#pragma unroll
  for (int i = 0; i < DX; ++i) a.x_final[(size_t)id * DX + i] = xu[i];
  if (Env::HAS_TERM) {
    double zt[DZT];
    Env::obs_term(xu, 0, ctx, zt);
#pragma unroll
    for (int i = 0; i < DZT; ++i) a.z_term[(size_t)id * DZT + i] = zt[i];
  }
}

template <class Env>
static int launch_rollout_t(const RolloutArgs& a, cudaStream_t s) {
  const long long n = (long long)a.B * a.R;
  rollout_kernel<Env><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace i2c
