// Sub-warp-per-problem variant of the EM sweep: G lanes of a warp cooperate on ONE problem (32 / G problems per warp).
//
// Why.  em_kernel / em_team_kernel give every problem one thread.  That is the right shape when the batch fills the
// machine (>= ~16k problems), but a tile of 32 problems is then ONE instruction stream: a double cart-pole cell is
// ~25 000 warp instructions (n = 7: the matrices do not fit the register file, 5.7 KB of spills per thread), i.e.
// ~60 000 cycles per cell on one sub-partition while 3/4 of the SM and every SM beyond the tile count idle.  BASELINE
// configs 4 and 5 shard to 2048 double cart-pole problems and 8192 quadrotor roll-outs per GPU -- exactly that regime.
// Here the small dense algebra of a cell is spread over G lanes: sigma points are evaluated one (pair) per lane, the
// right-hand sides of the triangular solves and the rows of the symmetric updates are distributed over the lanes, the
// Cholesky factorisations run column by column with the rows below the pivot in parallel.  Matrices live in shared
// memory (a few KB per problem), exchanged with __syncwarp() only: warps never synchronise with each other.
//
// Same arithmetic as Worker (i2c_kernels.cuh; reference lines cited there): identical formulas and, where the
// operation order is fixed by a dependency chain, identical order; sums that are reduced across lanes (cost / alpha
// statistics) differ by round-off.  Same records, flags, metrics, status words: the variants are interchangeable
// between launches.  Cubature inference only (Linearize / Gauss-Hermite / Riccati stay on the per-thread kernels).
#pragma once
#include <cstdio>

#include "i2c_kernels.cuh"

namespace i2c {

template <class Env>
struct GroupLay {
  using LY = Lay<Env>;
  static constexpr int DX = LY::DX, DU = LY::DU, N = LY::N, DZ = LY::DZ, DZT = LY::DZT;
  static constexpr int DYM = DZ > DX ? (DZ > DZT ? DZ : DZT) : (DX > DZT ? DX : DZT);
  static constexpr int P = 2 * N + 1;
  static constexpr int E_REC = LY::E_POST > LY::E_FILT ? LY::E_POST : LY::E_FILT;
  // per-problem shared-memory region (doubles); 2-D arrays are full row-major
  static constexpr int O_REC = 0;
  static constexpr int O_M0 = O_REC + 2 * E_REC; // (REC is double buffered)  carry mean [DX]
  static constexpr int O_S0 = O_M0 + DX;         // carry covariance [DX][DX] (full symmetric)
  static constexpr int O_L0 = O_S0 + DX * DX;    // its Cholesky factor (lower)
  static constexpr int O_I0 = O_L0 + DX * DX;    // reciprocal pivots [DX]
  static constexpr int O_MU = O_I0 + DX;         // joint mean [N]
  static constexpr int O_SIG = O_MU + N;         // joint covariance [N][N] (full symmetric)
  static constexpr int O_L = O_SIG + N * N;      // factor (lower)
  static constexpr int O_INVD = O_L + N * N;     // [N]
  static constexpr int O_Y = O_INVD + N;         // sigma-point images [P][DYM]
  static constexpr int O_MY = O_Y + P * DYM;     // [DYM]
  static constexpr int O_SYY = O_MY + DYM;       // [DYM][DYM] lower (in place: its factor)
  static constexpr int O_INVZ = O_SYY + DYM * DYM;
  static constexpr int O_RR = O_INVZ + DYM;      // residual / extra right-hand side [DYM]
  static constexpr int O_SXY = O_RR + DYM;       // [N][DYM]
  static constexpr int O_M3 = O_SXY + N * DYM;   // smoothed message mean [DX]
  static constexpr int O_S3 = O_M3 + DX;         // smoothed message covariance [DX][DX] full
  static constexpr int O_TMP = O_S3 + DX * DX;   // scratch [2 DX DX + 2 DX]
  static constexpr int SIZE = ((O_TMP + 2 * DX * DX + 2 * DX) + 1) & ~1;
};

// environment maps as functor types: one instantiation of the transform per (map, dimension), shared by the sweeps
template <class Env>
struct EvalObs {
  __device__ __forceinline__ void operator()(const double* x, int j, const typename Env::TrigT& c, const double*, double* y) const { Env::obs(x, j, c, y); }
};
template <class Env>
struct EvalObsTerm {
  __device__ __forceinline__ void operator()(const double* x, int j, const typename Env::TrigT& c, const double*, double* y) const { Env::obs_term(x, j, c, y); }
};
template <class Env>
struct EvalDyn {
  __device__ __forceinline__ void operator()(const double* x, int j, const typename Env::TrigT& c, const double* par, double* y) const { Env::dyn(x, j, c, par, y); }
};

#ifdef I2C_GROUP_TIMING
#define I2C_TICK(k)                                   \
  {                                                   \
    const long long now_ = clock64();                 \
    tacc[k] += now_ - tlast;                          \
    tlast = now_;                                     \
  }
#else
#define I2C_TICK(k)
#endif

template <class Env, int G>
struct GroupWorker {
#ifdef I2C_GROUP_TIMING
  long long tacc[16] = {0}, tlast = 0;
#endif
  using LY = Lay<Env>;
  using GL = GroupLay<Env>;
  using WK = Worker<Env, true>;
  using TrigT = typename Env::TrigT;
  static constexpr int DX = LY::DX, DU = LY::DU, N = LY::N, DZ = LY::DZ, DZT = LY::DZT, DYM = GL::DYM;
  static constexpr unsigned FULL = 0xffffffffu;

  WK wk;           // per-problem bookkeeping (record addressing, flags, targets, status, metrics): lane = problem slot
  const KParams& p;
  const int r;     // role of this lane inside its group
  double* sm;      // this problem's shared-memory region

  __device__ GroupWorker(const KParams& p_, int tile, int pl, int r_, double* sm_) : wk(p_, tile, pl, nullptr, nullptr), p(p_), r(r_), sm(sm_) {}

  __device__ __forceinline__ static void gsync() { __syncwarp(); }
  __device__ __forceinline__ static double gsum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
    return v;
  }
  __device__ __forceinline__ static double gmin(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o, G));
    return v;
  }

  // ---- cooperative primitives (every one ends synchronised) ---------------------------------------------------
  // Rows of a matrix have a FIXED owner lane (row i -> lane i mod G), every dimension is a template constant and
  // every loop is unrolled: shared-memory operands get compile-time offsets from one per-lane base pointer and the
  // independent dot-product chains of a lane interleave (the non-unrolled version spent 30 % of its instructions on
  // loop control and stalled on every 8-cycle DFMA dependency).
#define I2C_FOR_ROWS(i, n) _Pragma("unroll") for (int i##_0 = 0; i##_0 < (n); i##_0 += G) if (const int i = i##_0 + r; i < (n))

  // record of this problem: global [E][32] (pointer already at the problem's lane) -> shared REC buffer `buf`,
  // asynchronously (cp.async, 8 B per element); rec_wait() before the first read
  __device__ __forceinline__ void prefetch_rec(const double* g, int E, int buf) {
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sm + GL::O_REC + buf * GL::E_REC);
    for (int e = r; e < E; e += G)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + e * 8), "l"(g + (size_t)e * TILE) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void rec_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    gsync();
  }
  // Cholesky of the lower triangle of A (leading dimension LD) by columns J0 .. NN-1, rows below the pivot in parallel.
  // Rows < J0 already hold L (invd their reciprocal pivots) and the columns < J0 of the rows >= J0 have been eliminated
  // by the caller (build_joint does it for the one or two action rows without any barrier).  Two columns are retired per
  // barrier: every lane forms both pivots and the sub-diagonal entry L[j+1][j] redundantly, then updates its own rows in
  // both columns.  The diagonal of L is written once at the end.  Same operation order per entry as chol_rows.
  // (Tried and rejected: every lane factorising the whole matrix in registers -- the unrolled code of the five
  // factorisations per cell no longer fits the instruction cache: 6.0k instead of 2.4k cycles for the 7x7 factorisation.)
  template <int NN, int LD, int J0>
  __device__ __noinline__ static bool chol_cols_impl(double* A, double* invd, const int r) {
    // (static + explicit arguments: a non-inlined MEMBER function would take `this`, forcing the whole worker object
    // into local memory and every shared-memory access through generic loads)
    __builtin_assume(__isShared(A));
    __builtin_assume(__isShared(invd));
    bool ok = true;
    double mydiag[(NN + G - 1) / G], myinv[(NN + G - 1) / G], mysub[(NN + G - 1) / G];
#pragma unroll
    for (int j = J0; j < NN; j += 2) {
      const bool two = j + 1 < NN;
      // pivot j, entry (j+1, j), pivot j+1 -- redundantly in every lane
      double d0 = A[j * LD + j];
#pragma unroll
      for (int k = 0; k < j; ++k) d0 = fma(-A[j * LD + k], A[j * LD + k], d0);
      ok = ok && (d0 > 0.0) && (d0 < kFm[20]);
      const double r0 = fast_rsqrt(d0);
      double l10 = 0.0, d1 = 1.0, r1 = 1.0;
      if (two) {
        l10 = A[(j + 1) * LD + j];
#pragma unroll
        for (int k = 0; k < j; ++k) l10 = fma(-A[(j + 1) * LD + k], A[j * LD + k], l10);
        l10 *= r0;
        d1 = A[(j + 1) * LD + j + 1];
#pragma unroll
        for (int k = 0; k < j; ++k) d1 = fma(-A[(j + 1) * LD + k], A[(j + 1) * LD + k], d1);
        d1 = fma(-l10, l10, d1);
        ok = ok && (d1 > 0.0) && (d1 < kFm[20]);
        r1 = fast_rsqrt(d1);
      }
      I2C_FOR_ROWS(i, NN) {
        if (i > j + 1 || (!two && i > j)) {
          double s0 = A[i * LD + j], s1 = two ? A[i * LD + j + 1] : 0.0;
#pragma unroll
          for (int k = 0; k < j; ++k) {
            const double aik = A[i * LD + k];
            s0 = fma(-aik, A[j * LD + k], s0);
            if (two) s1 = fma(-aik, A[(j + 1) * LD + k], s1);
          }
          s0 *= r0;
          A[i * LD + j] = s0;
          if (two) A[i * LD + j + 1] = fma(-s0, l10, s1) * r1;
        } else if (i == j) {
          mydiag[i_0 / G] = d0 * r0;
          myinv[i_0 / G] = r0;
        } else if (two && i == j + 1) {
          mysub[i_0 / G] = l10;  // (stored at the end: the other lanes still read the raw entry in this phase, and
                                 //  nobody reads L[j+1][j] again inside this function)
          mydiag[i_0 / G] = d1 * r1;
          myinv[i_0 / G] = r1;
        }
      }
      gsync();
    }
    I2C_FOR_ROWS(i, NN) {
      if (i >= J0) {
        A[i * LD + i] = mydiag[i_0 / G];
        invd[i] = myinv[i_0 / G];
        if ((i - J0) & 1) A[i * LD + i - 1] = mysub[i_0 / G];
      }
    }
    gsync();
    return ok;
  }
  template <int NN, int LD, int J0>
  __device__ __forceinline__ bool chol_cols(double* A, double* invd) const {
    return chol_cols_impl<NN, LD, J0>(A, invd, r);
  }
  // y <- L^-1 y / L^-T y for a register vector
  template <int NN, int LD>
  __device__ __forceinline__ static void fsub(const double* L, const double* invd, double* y) {
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      double s = y[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-L[i * LD + k], y[k], s);
      y[i] = s * invd[i];
    }
  }
  template <int NN, int LD>
  __device__ __forceinline__ static void bsub(const double* L, const double* invd, double* y) {
#pragma unroll
    for (int i = NN - 1; i >= 0; --i) {
      double s = y[i];
#pragma unroll
      for (int k = i + 1; k < NN; ++k) s = fma(-L[k * LD + i], y[k], s);
      y[i] = s * invd[i];
    }
  }

  // Sigma-point transform around (m, L) in dimension D -> my[DY], Syy (lower, ld LDS, optionally + noise * Nz),
  // Sxy[D][DYM] = L * Dm.  Points +-j are spread over the lanes; then feature rows / state rows are.
  template <int D, int DY, int LDL, int LDS, class Eval>
  __device__ __forceinline__ void transform(const double* m, const double* L, double sf, double w0, double wi, Eval eval,
                                            double* my, double* Syy, double* Sxy, bool want_sxy, double noise, const double* Nz) {
    transform_impl<D, DY, LDL, LDS, Eval>(m, L, sf, w0, wi, wk.par, sm + GL::O_Y, my, Syy, Sxy, want_sxy, r);
    if (Nz) {  // observation noise on the rows this lane owns (it wrote them itself: no barrier needed)
      I2C_FOR_ROWS(a, DY) {
#pragma unroll
        for (int b = 0; b < DY; ++b)
          if (b <= a) Syy[a * LDS + b] = fma(noise, Nz[a * DY + b], Syy[a * LDS + b]);
      }
      gsync();
    }
  }
  template <int D, int DY, int LDL, int LDS, class Eval>
  __device__ __noinline__ static void transform_impl(const double* m, const double* L, double sf, double w0, double wi,
                                                     const double* par, double* Y, double* my, double* Syy, double* Sxy,
                                                     bool want_sxy, const int r) {
    __builtin_assume(__isShared(m));
    __builtin_assume(__isShared(L));
    __builtin_assume(__isShared(Y));
    __builtin_assume(__isShared(my));
    __builtin_assume(__isShared(Syy));
    if (want_sxy) __builtin_assume(__isShared(Sxy));
    const Eval eval;
    TrigT ctx;
    Env::center(m, ctx);
    const bool centre = w0 != 0.0;
    for (int pt = r; pt < 2 * D + (centre ? 1 : 0); pt += G) {  // (rolled: one copy of the map per instantiation)
      const int j = pt < 2 * D ? (pt >> 1) : -1;
      const bool minus = pt & 1;
      double x[D], y[DY];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        x[i] = m[i];
        if (j >= 0 && i >= j) {
          const double d = sf * L[i * LDL + j];
          x[i] = minus ? m[i] - d : m[i] + d;
        }
      }
      eval(x, j, ctx, par, y);
#pragma unroll
      for (int a = 0; a < DY; ++a) Y[pt * DYM + a] = y[a];
    }
    gsync();
    // every lane forms the full mean (DY * 2D adds, no exchange needed), the owner of row a stores it
    double mya[DY];
#pragma unroll
    for (int a = 0; a < DY; ++a) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) s += Y[(2 * j) * DYM + a] + Y[(2 * j + 1) * DYM + a];
      s = wi * s;
      if (centre) s = fma(w0, Y[(2 * D) * DYM + a], s);
      mya[a] = s;
    }
    I2C_FOR_ROWS(a, DY) {
      double mine = 0.0;
#pragma unroll
      for (int aa = 0; aa < DY; ++aa)
        if (aa == a) mine = mya[aa];
      my[a] = mine;
#pragma unroll
      for (int b = 0; b < DY; ++b) {
        if (b <= a) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double* yp = Y + (2 * j) * DYM;
            const double* ym = yp + DYM;
            s = fma(yp[a], yp[b], fma(ym[a], ym[b], s));
          }
          s = wi * s;
          if (centre) s = fma(w0 * Y[(2 * D) * DYM + a], Y[(2 * D) * DYM + b], s);
          s = fma(-mine, mya[b], s);
          Syy[a * LDS + b] = s;
        }
      }
    }
    if (want_sxy) {
      const double wsf = wi * sf;
      I2C_FOR_ROWS(i, D) {
#pragma unroll
        for (int a = 0; a < DY; ++a) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j)
            if (j <= i) s = fma(L[i * LDL + j], wsf * (Y[(2 * j) * DYM + a] - Y[(2 * j + 1) * DYM + a]), s);
          Sxy[i * DYM + a] = s;
        }
      }
    }
    gsync();
  }

  // Structured moments of the cost-feature maps (same shortcut as structured_obs_moments of the per-thread kernels;
  // valid for the cubature rule with zero centre weight): identity features have exact moments (copies of mu / Sigma),
  // only the NL trigonometric features are integrated, and only the sigma-point columns j <= OBS_JMAX move them.
  // Sig: the covariance the factor L belongs to (full, leading dimension LDL as L).
  template <int D, int DY, int LDL, bool TERM>
  __device__ __noinline__ static void structured_impl(const double* m, const double* Sig, const double* L, double sf, double wi,
                                                      double* Y, double* my, double* Syy, double* Sxy, double* Cx,
                                                      bool want_sxy, const int r) {
    __builtin_assume(__isShared(m));
    __builtin_assume(__isShared(Sig));
    __builtin_assume(__isShared(L));
    __builtin_assume(__isShared(Y));
    __builtin_assume(__isShared(my));
    __builtin_assume(__isShared(Syy));
    __builtin_assume(__isShared(Cx));
    if (want_sxy) __builtin_assume(__isShared(Sxy));
    constexpr int NL = Env::OBS_NL, JM = Env::OBS_JMAX, NLs = NL > 0 ? NL : 1, NP = 2 * (JM + 1);
    double mnl[NLs], Snl[TRI(NLs)];
    if constexpr (NL > 0) {
      double mc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) mc[i] = m[i];
      TrigT ctx;
      Env::center(mc, ctx);
      for (int pt = r; pt < NP; pt += G) {
        const int j = pt >> 1;
        const bool minus = pt & 1;
        double x[D], y[NL];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          x[i] = mc[i];
          if (i >= j) {
            const double d = sf * L[i * LDL + j];
            x[i] = minus ? mc[i] - d : mc[i] + d;
          }
        }
        Env::trig_nl(x, j, ctx, y);
#pragma unroll
        for (int a = 0; a < NL; ++a) Y[pt * DYM + a] = y[a];
      }
      gsync();
      double yc[NL], sy[NL], syy[TRI(NL)];
      Env::trig_nl(mc, -1, ctx, yc);
      constexpr double mult = 2.0 * (D - 1 - JM);
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        sy[a] = mult * yc[a];
#pragma unroll
        for (int b = 0; b <= a; ++b) syy[tix(a, b)] = mult * yc[a] * yc[b];
      }
#pragma unroll
      for (int j = 0; j <= JM; ++j) {
        const double* yp = Y + (2 * j) * DYM;
        const double* ym = yp + DYM;
#pragma unroll
        for (int a = 0; a < NL; ++a) {
          sy[a] += yp[a] + ym[a];
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
      // cross covariance of the state rows with the nonlinear block: Cx[i][k] = sum_{j <= min(i, JM)} L[i][j] Dm[j][k]
      const double wsf = wi * sf;
      I2C_FOR_ROWS(i, D) {
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          double v = 0.0;
#pragma unroll
          for (int j = 0; j <= JM; ++j)
            if (j <= i) v = fma(L[i * LDL + j], wsf * (Y[(2 * j) * DYM + k] - Y[(2 * j + 1) * DYM + k]), v);
          Cx[i * NLs + k] = v;
        }
      }
      gsync();
    }
    // assemble: feature row a per lane (my, Syy lower), state row i per lane (Sxy)
    I2C_FOR_ROWS(a, DY) {
      const int sa = TERM ? Env::term_src(a) : Env::obs_src(a);
      double v = 0.0;
      if (sa >= 0) {
        v = m[sa];
      } else {
#pragma unroll
        for (int k = 0; k < NLs; ++k)
          if (k == -1 - sa) v = mnl[k];
      }
      my[a] = v;
#pragma unroll
      for (int b = 0; b < DY; ++b) {
        if (b <= a) {
          const int sb = TERM ? Env::term_src(b) : Env::obs_src(b);  // (b is a compile-time constant here)
          double w = 0.0;
          if (sa >= 0 && sb >= 0) {
            w = Sig[sa * LDL + sb];
          } else if (sa < 0 && sb < 0) {
#pragma unroll
            for (int k = 0; k < NLs; ++k)
#pragma unroll
              for (int l = 0; l < NLs; ++l)
                if (k == -1 - sa && l == -1 - sb) w = Snl[six(k, l)];
          } else if (sa < 0) {
            w = Cx[(sb >= 0 ? sb : 0) * NLs + (-1 - sa)];
          } else {
            w = Cx[sa * NLs + (sb < 0 ? -1 - sb : 0)];
          }
          Syy[a * DYM + b] = w;
        }
      }
    }
    if (want_sxy) {
      I2C_FOR_ROWS(i, D) {
#pragma unroll
        for (int a = 0; a < DY; ++a) {
          const int sa = TERM ? Env::term_src(a) : Env::obs_src(a);
          Sxy[i * DYM + a] = sa >= 0 ? Sig[i * LDL + (sa >= 0 ? sa : 0)] : Cx[i * NLs + (sa < 0 ? -1 - sa : 0)];
        }
      }
    }
    gsync();
  }
  // cost-feature moments of (m, Sig, L): structured when the rule allows it, else the generic sigma-point transform
  template <int D, int DY, int LDL, bool TERM>
  __device__ __forceinline__ void obs_moments(const double* m, const double* Sig, const double* L, double sf, double w0, double wi,
                                              bool want_sxy, double noise, const double* Nz) {
    double* my = sm + GL::O_MY;
    double* Syy = sm + GL::O_SYY;
    double* Sxy = sm + GL::O_SXY;
    if (p.fast_obs) {
      structured_impl<D, DY, LDL, TERM>(m, Sig, L, sf, wi, sm + GL::O_Y, my, Syy, Sxy, sm + GL::O_TMP, want_sxy, r);
      if (Nz) {
        I2C_FOR_ROWS(a, DY) {
#pragma unroll
          for (int b = 0; b < DY; ++b)
            if (b <= a) Syy[a * DYM + b] = fma(noise, Nz[a * DY + b], Syy[a * DYM + b]);
        }
        gsync();
      }
    } else if constexpr (TERM) {
      transform<D, DY, LDL, DYM>(m, L, sf, w0, wi, EvalObsTerm<Env>(), my, Syy, Sxy, want_sxy, noise, Nz);
    } else {
      transform<D, DY, LDL, DYM>(m, L, sf, w0, wi, EvalObs<Env>(), my, Syy, Sxy, want_sxy, noise, Nz);
    }
  }

  // Gaussian conditioning of (mu[D], Sig[D][D] full, LD) on the observation whose moments sit in MY / SYY (noise added) /
  // SXY, with target z: W = Lz^-1 Sxy^T, mu += W^T Lz^-1 (z - my), Sig -= W^T W.
  template <int D, int DY, int LD>
  __device__ __forceinline__ bool condition(double* mu, double* Sig, const double* z) {
    // the residual z - my is formed here (z lives in registers), the rest in the shared non-inlined body
    double* rz = sm + GL::O_RR;
#pragma unroll
    for (int a = 0; a < DY; ++a)
      if ((a % G) == r) rz[a] = z[a] - sm[GL::O_MY + a];
    return condition_impl<D, DY, LD>(mu, Sig, rz, sm + GL::O_SYY, sm + GL::O_INVZ, sm + GL::O_SXY, r);
  }
  template <int D, int DY, int LD>
  __device__ __noinline__ static bool condition_impl(double* mu, double* Sig, const double* rz, double* Sz, double* invz, double* W,
                                                     const int r) {
    __builtin_assume(__isShared(mu));
    __builtin_assume(__isShared(Sig));
    __builtin_assume(__isShared(rz));
    __builtin_assume(__isShared(Sz));
    __builtin_assume(__isShared(invz));
    __builtin_assume(__isShared(W));
    const bool ok = chol_cols_impl<DY, DYM, 0>(Sz, invz, r);  // (its barriers also publish rz)
    // residual solve: every lane (DY^2 / 2 FMAs, no exchange)
    double rr[DY];
#pragma unroll
    for (int a = 0; a < DY; ++a) rr[a] = rz[a];
    fsub<DY, DYM>(Sz, invz, rr);
    I2C_FOR_ROWS(c, D) {
      double w[DY];
#pragma unroll
      for (int a = 0; a < DY; ++a) w[a] = W[c * DYM + a];
      fsub<DY, DYM>(Sz, invz, w);
      double dm = 0.0;
#pragma unroll
      for (int a = 0; a < DY; ++a) {
        W[c * DYM + a] = w[a];
        dm = fma(w[a], rr[a], dm);
      }
      mu[c] += dm;
    }
    gsync();
    I2C_FOR_ROWS(i, D) {
      double wi_[DY];
#pragma unroll
      for (int a = 0; a < DY; ++a) wi_[a] = W[i * DYM + a];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (j <= i) {
          double s = Sig[i * LD + j];
#pragma unroll
          for (int a = 0; a < DY; ++a) s = fma(-wi_[a], W[j * DYM + a], s);
          Sig[i * LD + j] = s;
          Sig[j * LD + i] = s;
        }
      }
    }
    gsync();
    return ok;
  }

  // quadratic-cost statistics / alpha trace of the feature moments in MY / SYY (lower), reduced over the group
  template <int DZ_>
  __device__ __forceinline__ void cost_stats(const double* zref, double& mean, double& var) {
    const double* mz = sm + GL::O_MY;
    const double* Sz = sm + GL::O_SYY;
    double m = 0.0, t2 = 0.0, q4 = 0.0;
    if (p.qr_diag) {
      I2C_FOR_ROWS(a, DZ_) {
        const double qa = p.QR[a * DZ_ + a], ea = mz[a] - zref[a], va = qa * ea;
        m = fma(ea, va, m);
        m = fma(Sz[a * DYM + a], qa, m);
#pragma unroll
        for (int b = 0; b < DZ_; ++b) {
          const double s = a >= b ? Sz[a * DYM + b] : Sz[b * DYM + a];
          const double vb = p.QR[b * DZ_ + b] * (mz[b] - zref[b]);
          t2 = fma(s * s, qa * p.QR[b * DZ_ + b], t2);
          q4 = fma(va * s, vb, q4);
        }
      }
    } else {
      // P = Sz QR (row a per lane, parked in the SXY area), v = QR e
      double* P = sm + GL::O_SXY;
      double* v = sm + GL::O_RR;
      I2C_FOR_ROWS(a, DZ_) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < DZ_; ++b) s = fma(p.QR[a * DZ_ + b], mz[b] - zref[b], s);
        v[a] = s;
#pragma unroll
        for (int b = 0; b < DZ_; ++b) {
          double t = 0.0;
#pragma unroll
          for (int c = 0; c < DZ_; ++c) t = fma(a >= c ? Sz[a * DYM + c] : Sz[c * DYM + a], p.QR[c * DZ_ + b], t);
          P[a * DYM + b] = t;
        }
      }
      gsync();
      I2C_FOR_ROWS(a, DZ_) {
        m = fma(mz[a] - zref[a], v[a], m);
        m += P[a * DYM + a];
#pragma unroll
        for (int b = 0; b < DZ_; ++b) {
          t2 = fma(P[a * DYM + b], P[b * DYM + a], t2);
          q4 = fma(v[a] * (a >= b ? Sz[a * DYM + b] : Sz[b * DYM + a]), v[b], q4);
        }
      }
      gsync();
    }
    mean = gsum(m);
    var = 2.0 * gsum(t2) + 4.0 * gsum(q4);
  }
  template <int DZ_>
  __device__ __forceinline__ double alpha_trace(const double* Q, int qr_diag, const double* z) {
    const double* mz = sm + GL::O_MY;
    const double* Sz = sm + GL::O_SYY;
    double tr = 0.0;
    I2C_FOR_ROWS(a, DZ_) {
      const double da = z[a] - mz[a];
      if (qr_diag) {
        tr = fma(Q[a * DZ_ + a], fma(da, da, Sz[a * DYM + a]), tr);
      } else {
#pragma unroll
        for (int b = 0; b < DZ_; ++b)
          tr = fma(Q[a * DZ_ + b], fma(z[b] - mz[b], da, a >= b ? Sz[a * DYM + b] : Sz[b * DYM + a]), tr);
      }
    }
    return gsum(tr);
  }

  // exp(-1/2 d^T C^-1 d), C = Sxx(prior record) + S0, d = m0 - mx   (i2c.py:369-374); rho identical in every lane
  __device__ __forceinline__ bool pdf_ratio(const double* rec, double& rho) {
    const double* S0 = sm + GL::O_S0;
    const double* m0 = sm + GL::O_M0;
    double* C = sm + GL::O_TMP;
    double* ic = C + DX * DX;
    I2C_FOR_ROWS(i, DX) {
#pragma unroll
      for (int j = 0; j < DX; ++j)
        if (j <= i) C[i * DX + j] = rec[LY::P_SIG + tix(i, j)] + S0[i * DX + j];
    }
    gsync();
    const bool ok = chol_cols<DX, DX, 0>(C, ic);
    double d[DX], q = 0.0;
#pragma unroll
    for (int i = 0; i < DX; ++i) d[i] = m0[i] - rec[LY::P_MU + i];
    fsub<DX, DX>(C, ic, d);
#pragma unroll
    for (int i = 0; i < DX; ++i) q = fma(d[i], d[i], q);
    rho = exp(-0.5 * q);
    return ok;
  }

  // joint (x,u) around the carried message under the controller Kt (registers, identical in every lane):
  // MU = [m0; mu_u], SIG = [[S0, (Kt S0)^T], [Kt S0, Suu]], L = chol (rows < DX from the carried factor)
  __device__ __forceinline__ bool build_joint(const double* Kt, const double* mu_u, const double* Suu, bool coupled) {
    const double* S0 = sm + GL::O_S0;
    const double* L0 = sm + GL::O_L0;
    const double* I0 = sm + GL::O_I0;
    double* MU = sm + GL::O_MU;
    double* SIG = sm + GL::O_SIG;
    double* L = sm + GL::O_L;
    double* invd = sm + GL::O_INVD;
    I2C_FOR_ROWS(i, N) {
      if (i < DX) {
        MU[i] = sm[GL::O_M0 + i];
        invd[i] = I0[i];
#pragma unroll
        for (int j = 0; j < DX; ++j) {
          SIG[i * N + j] = S0[i * DX + j];
          if (j <= i) L[i * N + j] = L0[i * DX + j];
        }
      } else {
#pragma unroll
        for (int u = 0; u < DU; ++u) {
          if (u == i - DX) {
            MU[i] = mu_u[u];
            double row[DX];
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              double s = 0.0;
              if (coupled) {
#pragma unroll
                for (int k = 0; k < DX; ++k) s = fma(Kt[u * DX + k], S0[k * DX + j], s);
              }
              SIG[i * N + j] = s;
              SIG[j * N + i] = s;
              row[j] = s;
            }
#pragma unroll
            for (int q = 0; q < DU; ++q) SIG[i * N + DX + q] = Suu[u >= q ? tix(u, q) : tix(q, u)];
            // columns < DX of this row against the carried factor: L[i][j] = (Sig[i][j] - sum_k L[i][k] L0[j][k]) / L0[j][j]
#pragma unroll
            for (int j = 0; j < DX; ++j) {
              double s = row[j];
#pragma unroll
              for (int k = 0; k < j; ++k) s = fma(-row[k], L0[j * DX + k], s);
              row[j] = s * I0[j];
              L[i * N + j] = row[j];
            }
#pragma unroll
            for (int q = 0; q <= u; ++q) L[i * N + DX + q] = Suu[tix(u, q)];
          }
        }
      }
    }
    gsync();
    return chol_cols<N, N, DX>(L, invd);
  }

  // store helpers: lanes share the elements of a vector / packed-lower matrix held in shared memory
  template <int NN>
  __device__ __forceinline__ void put_vec(double* g, int off, const double* v) {
    I2C_FOR_ROWS(i, NN) g[(size_t)(off + i) * TILE] = v[i];
  }
  template <int NN, int LD>
  __device__ __forceinline__ void put_tri(double* g, int off, const double* A) {
    I2C_FOR_ROWS(i, NN) {
#pragma unroll
      for (int j = 0; j < NN; ++j)
        if (j <= i) g[(size_t)(off + tix(i, j)) * TILE] = A[i * LD + j];
    }
  }

  // ---------------------------------------------------------------------------------- forward cell (i2c.py:350-447)
  // rec: this cell's prior record (already waited for)
  __device__ void forward_cell(int it, int t, int flags, double alpha, bool aux, const double* rec, LogAcc& ent_x) {
    double* MU = sm + GL::O_MU;
    double* SIG = sm + GL::O_SIG;
    double* L = sm + GL::O_L;
    double* invd = sm + GL::O_INVD;
    {
      double mu_u[DU], Suu[TRI(DU)], Kt[DU * DX];
#pragma unroll
      for (int u = 0; u < DU; ++u) mu_u[u] = rec[LY::P_MU + DX + u];
#pragma unroll
      for (int u = 0; u < DU; ++u)
#pragma unroll
        for (int q = 0; q <= u; ++q) Suu[tix(u, q)] = rec[LY::P_SIG + tix(DX + u, DX + q)];
      const bool indep = flags & I2C_CELL_INDEPENDENT;
      if (!indep) {
        double rho = 1.0;
        if (!pdf_ratio(rec, rho)) wk.fail(I2C_FAIL_MVN, it, t);
        const double* S0 = sm + GL::O_S0;
        double d[DX], KS[DU * DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) d[i] = sm[GL::O_M0 + i] - rec[LY::P_MU + i];
#pragma unroll
        for (int i = 0; i < DU * DX; ++i) Kt[i] = rec[LY::P_K + i] * rho;
#pragma unroll
        for (int u = 0; u < DU; ++u)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(Kt[u * DX + k], S0[k * DX + j], s);
            KS[u * DX + j] = s;
          }
#pragma unroll
        for (int u = 0; u < DU; ++u) {
          double s = mu_u[u];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(Kt[u * DX + k], d[k], s);
          mu_u[u] = s;
#pragma unroll
          for (int q = 0; q <= u; ++q) {
            double v = Suu[tix(u, q)];
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(-Kt[u * DX + k], rec[LY::P_SIG + tix(DX + q, k)], v);
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(KS[u * DX + k], Kt[q * DX + k], v);
            Suu[tix(u, q)] = v;
          }
        }
      }
      if (!build_joint(Kt, mu_u, Suu, !indep)) wk.fail(I2C_FAIL_CHOL_PRIOR, it, t);
    I2C_TICK(0)
    }
    double* af = aux ? wk.rec(p.auxf, t, LY::E_AUXF) : nullptr;
    if (aux) {
      put_vec<N>(af, LY::AF_MU0, MU);
      put_tri<N, N>(af, LY::AF_SIG0, SIG);
    }
    // ---- cost observation update (i2c.py:390-404)
    {
      const double a_cell = wk.cell_alpha(t, flags, alpha);
      obs_moments<N, DZ, N, false>(MU, SIG, L, p.sf_n, p.w0_n, p.wi_n, true, a_cell, p.QRinv);
      if (aux) {
        put_vec<DZ>(af, LY::AF_MUZ, sm + GL::O_MY);
        put_tri<DZ, DYM>(af, LY::AF_SIGZ, sm + GL::O_SYY);
      }
      I2C_TICK(1)
      double z[DZ];
      wk.load_z(t, z);
      if (!condition<N, DZ, N>(MU, SIG, z)) wk.fail(I2C_FAIL_CHOL_OBS, it, t);
    I2C_TICK(2)
    }
    double* fr = wk.rec(p.filt, t, LY::E_FILT);
    put_vec<N>(fr, LY::F_MU1, MU);
    put_tri<N, N>(fr, LY::F_SIG1, SIG);
    // ---- dynamics moment matching (i2c.py:415-428)
    I2C_FOR_ROWS(i, N) {
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j <= i) L[i * N + j] = SIG[i * N + j];
    }
    gsync();
    if (!chol_cols<N, N, 0>(L, invd)) wk.fail(I2C_FAIL_CHOL_FILTERED, it, t);
    I2C_TICK(3)
    double* M0 = sm + GL::O_M0;
    double* S0 = sm + GL::O_S0;
    double* L0 = sm + GL::O_L0;
    double* I0 = sm + GL::O_I0;
    transform<N, DX, N, DX>(MU, L, p.sf_n, p.w0_n, p.wi_n,
                            EvalDyn<Env>(), M0, S0,
                            sm + GL::O_SXY, true, 0.0, nullptr);
    I2C_TICK(4)
    I2C_FOR_ROWS(i, DX) {
#pragma unroll
      for (int j = 0; j < DX; ++j)
        if (j <= i) {
          const double s = S0[i * DX + j] + p.sig_eta[tix(i, j)];
          S0[i * DX + j] = s;
          S0[j * DX + i] = s;
          L0[i * DX + j] = s;
        }
    }
    gsync();
    if (!chol_cols<DX, DX, 0>(L0, I0)) wk.fail(I2C_FAIL_CHOL_X3, it, t);
    I2C_TICK(5)
    // J_dyn = Sxy Sig_x3^-1: row i per lane
    I2C_FOR_ROWS(i, N) {
      double w[DX];
#pragma unroll
      for (int a = 0; a < DX; ++a) w[a] = sm[GL::O_SXY + i * DYM + a];
      fsub<DX, DX>(L0, I0, w);
      bsub<DX, DX>(L0, I0, w);
#pragma unroll
      for (int a = 0; a < DX; ++a) fr[(size_t)(LY::F_J + i * DX + a) * TILE] = w[a];
    }
    // ---- terminal cost update on the outgoing message (i2c.py:430-443)
    if (Env::HAS_TERM && (flags & I2C_CELL_TERMINAL) && p.has_qf) {
      const double a_cell = wk.cell_alpha(t, flags, alpha);
      gsync();  // the J rows above still read SXY
      obs_moments<DX, DZT, DX, true>(M0, S0, L0, p.sf_x, p.w0_x, p.wi_x, true, a_cell, p.Qfinv);
      double zt[DZT];
      wk.load_zterm(zt);
      if (!condition<DX, DZT, DX>(M0, S0, zt)) wk.fail(I2C_FAIL_CHOL_TERMINAL, it, t);
      I2C_FOR_ROWS(i, DX) {
#pragma unroll
        for (int j = 0; j < DX; ++j)
          if (j <= i) L0[i * DX + j] = S0[i * DX + j];
      }
      gsync();
      if (!chol_cols<DX, DX, 0>(L0, I0)) wk.fail(I2C_FAIL_CHOL_TERMINAL, it, t);
    }
    put_vec<DX>(fr, LY::F_MU3, M0);
    I2C_TICK(6)
    put_tri<DX, DX>(fr, LY::F_SIG3, S0);
#pragma unroll
    for (int i = 0; i < DX; ++i) ent_x.mul(L0[i * DX + i]);
  }

  // ---------------------------------------------------------------------------------- backward cell (i2c.py:544-610)
  struct Stats {
    double cost, cost_var, tr;
    LogAcc ent_u;
  };
  // rec: this cell's filtered record (already waited for)
  __device__ void backward_cell(int it, int t, bool aux, const double* rec, Stats& st) {
    double* MU = sm + GL::O_MU;
    double* SIG = sm + GL::O_SIG;
    double* L = sm + GL::O_L;
    double* invd = sm + GL::O_INVD;
    double* M3 = sm + GL::O_M3;
    double* S3 = sm + GL::O_S3;
    if (aux) {
      double* ab = wk.rec(p.auxb, t, LY::E_AUXB);
      put_vec<DX>(ab, LY::AB_MU3M, M3);
      put_tri<DX, DX>(ab, LY::AB_SIG3M, S3);
    }
    // dm, dS: every lane (DX^2 subtractions, no exchange)
    double dm[DX], dS[TRI(DX)];
#pragma unroll
    for (int i = 0; i < DX; ++i) dm[i] = M3[i] - rec[LY::F_MU3 + i];
#pragma unroll
    for (int i = 0; i < DX; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) dS[tix(i, j)] = S3[i * DX + j] - rec[LY::F_SIG3 + tix(i, j)];
    gsync();  // M3 / S3 read by everyone before their owners overwrite them below
    // mu_xu1_m = mu_xu1_f + J dm;  sig_xu1_m = sig_xu1_f + J dS J^T   (row i per lane)
    const double* J = rec + LY::F_J;
    I2C_FOR_ROWS(i, N) {
      double s = rec[LY::F_MU1 + i], JD[DX], Ji[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) Ji[k] = J[i * DX + k];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(Ji[k], dm[k], s);
      MU[i] = s;
#pragma unroll
      for (int k = 0; k < DX; ++k) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < DX; ++l) v = fma(Ji[l], dS[six(l, k)], v);
        JD[k] = v;
      }
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j <= i) {
          double v = rec[LY::F_SIG1 + tix(i, j)];
#pragma unroll
          for (int k = 0; k < DX; ++k) v = fma(JD[k], J[j * DX + k], v);
          SIG[i * N + j] = v;
          SIG[j * N + i] = v;
          L[i * N + j] = v;
          if (i < DX) {
            S3[i * DX + j] = v;
            S3[j * DX + i] = v;
          }
        }
      if (i < DX) M3[i] = s;
    }
    gsync();
    double* po = wk.rec(wk.post, t, LY::E_POST);
    put_vec<N>(po, LY::P_MU, MU);
    put_tri<N, N>(po, LY::P_SIG, SIG);
    I2C_TICK(7)
    // ---- everything that does not feed the recursion: controller, cost-feature moments, statistics
    if (!chol_cols<N, N, 0>(L, invd)) wk.fail(I2C_FAIL_CHOL_POSTERIOR, it, t);
    I2C_TICK(8)
    I2C_FOR_ROWS(u, DU) {
      double w[DX];
#pragma unroll
      for (int j = 0; j < DX; ++j) w[j] = L[(DX + u) * N + j];
      bsub<DX, N>(L, invd, w);
      double kk = MU[DX + u];
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        po[(size_t)(LY::P_K + u * DX + j) * TILE] = w[j];
        kk = fma(-w[j], MU[j], kk);
      }
      po[(size_t)(LY::P_KK + u) * TILE] = kk;
#pragma unroll
      for (int q = 0; q < DU; ++q)
        if (q <= u) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DU; ++k)
            if (k <= q) s = fma(L[(DX + u) * N + DX + k], L[(DX + q) * N + DX + k], s);
          po[(size_t)(LY::P_SIGK + tix(u, q)) * TILE] = s;
        }
    }
    {
      double Su[TRI(DU)], iu[DU];
#pragma unroll
      for (int u = 0; u < DU; ++u)
#pragma unroll
        for (int q = 0; q <= u; ++q) Su[tix(u, q)] = SIG[(DX + u) * N + DX + q];
      if (!chol_rows<DU>(Su, iu)) wk.fail(I2C_FAIL_POLICY_DET, it, t);
#pragma unroll
      for (int u = 0; u < DU; ++u) st.ent_u.mul(Su[tix(u, u)]);
    }
    I2C_TICK(9)
    obs_moments<N, DZ, N, false>(MU, SIG, L, p.sf_n, p.w0_n, p.wi_n, false, 0.0, nullptr);
    I2C_TICK(12)
    if (aux) {
      double* ab = wk.rec(p.auxb, t, LY::E_AUXB);
      put_vec<DZ>(ab, LY::AB_MUZ, sm + GL::O_MY);
      put_tri<DZ, DYM>(ab, LY::AB_SIGZ, sm + GL::O_SYY);
    }
    double cm, cv, z[DZ];
    cost_stats<DZ>(p.z_graph, cm, cv);
    I2C_TICK(10)
    st.cost += cm;
    st.cost_var += cv;
    wk.load_z(t, z);
    st.tr += alpha_trace<DZ>(p.QR, p.qr_diag, z);
    I2C_TICK(11)
  }

  // end of chain (i2c.py:546-572): covariance control or plain hand-over + terminal cost-feature moments
  __device__ void backward_terminal(int it, int t, double temp, double a_cell, double& tr_term) {
    const double* M0 = sm + GL::O_M0;
    const double* S0 = sm + GL::O_S0;
    const double* L0 = sm + GL::O_L0;
    const double* I0 = sm + GL::O_I0;
    double* M3 = sm + GL::O_M3;
    double* S3 = sm + GL::O_S3;
    double* Lm = sm + GL::O_L;  // factor of sig_x3_m for the terminal transform (ld N)
    double* im = sm + GL::O_INVD;
    gsync();
    if (p.cov_ctrl) {
      double* A = sm + GL::O_TMP;
      double* ia = A + DX * DX;
      double* W = sm + GL::O_SXY;  // W[j][:] = La^-1 S[:, j]   (ld DYM)
      I2C_FOR_ROWS(i, DX) {
#pragma unroll
        for (int j = 0; j < DX; ++j)
          if (j <= i) A[i * DX + j] = p.sxt[tix(i, j)] + temp * S0[i * DX + j];
      }
      gsync();
      if (!chol_cols<DX, DX, 0>(A, ia)) wk.fail(I2C_FAIL_COV_CONTROL, it, t);
      I2C_FOR_ROWS(j, DX) {
        double w[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) w[i] = temp * S0[i * DX + j];
        fsub<DX, DX>(A, ia, w);
#pragma unroll
        for (int i = 0; i < DX; ++i) W[j * DYM + i] = w[i];
      }
      gsync();
      I2C_FOR_ROWS(i, DX) {
#pragma unroll
        for (int j = 0; j < DX; ++j)
          if (j <= i) {
            double s = temp * S0[i * DX + j];
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(-W[i * DYM + k], W[j * DYM + k], s);
            S3[i * DX + j] = s;
            S3[j * DX + i] = s;
            Lm[i * N + j] = s;
          }
      }
      double v[DX];
#pragma unroll
      for (int i = 0; i < DX; ++i) v[i] = M0[i];
      fsub<DX, DX>(L0, I0, v);
      bsub<DX, DX>(L0, I0, v);
      const double it_ = 1.0 / temp;
#pragma unroll
      for (int i = 0; i < DX; ++i) v[i] = fma(v[i], it_, p.sxt_inv_mu[i]);
      gsync();
      I2C_FOR_ROWS(i, DX) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DX; ++k) s = fma(S3[i * DX + k], v[k], s);
        M3[i] = s;
      }
      gsync();
      if (Env::HAS_TERM && p.has_qf) {
        if (!chol_cols<DX, N, 0>(Lm, im)) wk.fail(I2C_FAIL_COV_CONTROL, it, t);
      }
    } else {
      I2C_FOR_ROWS(i, DX) {
        M3[i] = M0[i];
        im[i] = I0[i];
#pragma unroll
        for (int j = 0; j < DX; ++j) {
          S3[i * DX + j] = S0[i * DX + j];
          if (j <= i) Lm[i * N + j] = L0[i * DX + j];
        }
      }
      gsync();
    }
    tr_term = 0.0;
    if (Env::HAS_TERM && p.has_qf) {
      transform<DX, DZT, N, DYM>(M3, Lm, p.sf_x, p.w0_x, p.wi_x,
                                 EvalObsTerm<Env>(),
                                 sm + GL::O_MY, sm + GL::O_SYY, nullptr, false, 0.0, nullptr);
      double* tm = p.term + ((size_t)wk.tile * LY::E_TERM) * TILE + wk.lane;
      put_vec<DZT>(tm, LY::TM_MU, sm + GL::O_MY);
      put_tri<DZT, DYM>(tm, LY::TM_SIG, sm + GL::O_SYY);
      double zt[DZT];
      wk.load_zterm(zt);
      tr_term = alpha_trace<DZT>(p.Qf, 0, zt);
    }
  }

  // ---------------------------------------------------------------------------------- propagate cell (i2c.py:150-199)
  struct PStats {
    double cost, cost_var, cost_min, tr;
    LogAcc ent;
  };
  __device__ void propagate_cell(int it, int t, int flags, bool aux, const double* rec, PStats& st) {
    double* MU = sm + GL::O_MU;
    double* SIG = sm + GL::O_SIG;
    double* L = sm + GL::O_L;
    {
      double mu_u[DU], Suu[TRI(DU)], Kt[DU * DX];
#pragma unroll
      for (int u = 0; u < DU; ++u) mu_u[u] = rec[LY::P_MU + DX + u];
#pragma unroll
      for (int u = 0; u < DU; ++u)
#pragma unroll
        for (int q = 0; q <= u; ++q) Suu[tix(u, q)] = rec[LY::P_SIG + tix(DX + u, DX + q)];
#pragma unroll
      for (int i = 0; i < DU * DX; ++i) Kt[i] = rec[LY::P_K + i];
      if (!(flags & I2C_CELL_INDEPENDENT)) {
        const double* S0 = sm + GL::O_S0;
        double d[DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) d[i] = sm[GL::O_M0 + i] - rec[LY::P_MU + i];
        if (flags & I2C_CELL_EXPERT) {
          double rho;
          if (pdf_ratio(rec, rho)) {  // the reference swallows the exception: K unchanged (i2c.py:161-167)
#pragma unroll
            for (int i = 0; i < DU * DX; ++i) Kt[i] *= rho;
          }
        }
#pragma unroll
        for (int u = 0; u < DU; ++u) {
          double s = mu_u[u];
#pragma unroll
          for (int k = 0; k < DX; ++k) s = fma(Kt[u * DX + k], d[k], s);
          mu_u[u] = s;
        }
        double KD[DU * DX];
#pragma unroll
        for (int u = 0; u < DU; ++u)
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DX; ++k) s = fma(Kt[u * DX + k], S0[k * DX + j] - rec[LY::P_SIG + six(k, j)], s);
            KD[u * DX + j] = s;
          }
#pragma unroll
        for (int u = 0; u < DU; ++u)
#pragma unroll
          for (int q = 0; q <= u; ++q) {
            double v = Suu[tix(u, q)];
#pragma unroll
            for (int k = 0; k < DX; ++k) v = fma(KD[u * DX + k], Kt[q * DX + k], v);
            Suu[tix(u, q)] = v;
          }
      }
      if (!build_joint(Kt, mu_u, Suu, true)) wk.fail(I2C_FAIL_CHOL_PROPAGATE, it, t);
    }
    double* pf = aux ? wk.rec(p.pf, t, LY::E_PF) : nullptr;
    if (aux) {
      put_vec<N>(pf, LY::PF_MU, MU);
      put_tri<N, N>(pf, LY::PF_SIG, SIG);
    }
    obs_moments<N, DZ, N, false>(MU, SIG, L, p.sf_n, p.w0_n, p.wi_n, false, 0.0, nullptr);
    if (aux) {
      put_vec<DZ>(pf, LY::PF_MUZ, sm + GL::O_MY);
      put_tri<DZ, DYM>(pf, LY::PF_SIGZ, sm + GL::O_SYY);
    }
    double cm, cv, z[DZ];
    cost_stats<DZ>(p.z_graph, cm, cv);
    st.cost += cm;
    st.cost_var += cv;
    st.cost_min = fmin(st.cost_min, cm);
    wk.load_z(t, z);
    st.tr += alpha_trace<DZ>(p.QR, p.qr_diag, z);
    double* M0 = sm + GL::O_M0;
    double* S0 = sm + GL::O_S0;
    double* L0 = sm + GL::O_L0;
    double* I0 = sm + GL::O_I0;
    transform<N, DX, N, DX>(MU, L, p.sf_n, p.w0_n, p.wi_n,
                            EvalDyn<Env>(), M0, S0,
                            nullptr, false, 0.0, nullptr);
    I2C_FOR_ROWS(i, DX) {
#pragma unroll
      for (int j = 0; j < DX; ++j)
        if (j <= i) {
          const double s = S0[i * DX + j] + p.sig_eta[tix(i, j)];
          S0[i * DX + j] = s;
          S0[j * DX + i] = s;
          L0[i * DX + j] = s;
        }
    }
    gsync();
    if (!chol_cols<DX, DX, 0>(L0, I0)) wk.fail(I2C_FAIL_CHOL_PROPAGATE, it, t);
#pragma unroll
    for (int i = 0; i < DX; ++i) st.ent.mul(L0[i * DX + i]);
    if (aux) {
      put_vec<DX>(pf, LY::PF_MU3, M0);
      put_tri<DX, DX>(pf, LY::PF_SIG3, S0);
    }
  }

  // carry <- a Gaussian stored tiled in global memory (mean at gm, packed covariance at gS, element stride TILE)
  __device__ bool load_carry(const double* gm, const double* gS) {
    double* M0 = sm + GL::O_M0;
    double* S0 = sm + GL::O_S0;
    double* L0 = sm + GL::O_L0;
    gsync();
    I2C_FOR_ROWS(i, DX) {
      M0[i] = gm[(size_t)i * TILE];
#pragma unroll
      for (int j = 0; j < DX; ++j)
        if (j <= i) {
          const double s = gS[(size_t)tix(i, j) * TILE];
          S0[i * DX + j] = s;
          S0[j * DX + i] = s;
          L0[i * DX + j] = s;
        }
    }
    gsync();
    return chol_cols<DX, DX, 0>(L0, sm + GL::O_I0);
  }
  __device__ bool load_x0() {
    return load_carry(p.x0 + ((size_t)wk.tile * DX) * TILE + wk.lane, p.sig_x0 + ((size_t)wk.tile * TRI(DX)) * TILE + wk.lane);
  }
  __device__ __forceinline__ const double* rec_buf(int t) const { return sm + GL::O_REC + (t & 1) * GL::E_REC; }

  // ---------------------------------------------------------------------------------- the EM loop (Worker::run_impl<false>)
  __device__ void run() {
#ifdef I2C_GROUP_TIMING
    tlast = clock64();
#endif
    const double HALF_LOG_2PIE = 1.4189385332046727;
    double alpha = p.alpha[wk.b];
    const bool aux = p.phases & I2C_PH_STORE_AUX;
    const bool writer = r == 0;
    bool flipped = false;
    double temp = p.temp0;
    const int T = p.T;
    for (int it = 0; it < p.n_iter; ++it) {
      LogAcc ent_x;
      ent_x.reset();
      Stats st;
      st.cost = st.cost_var = st.tr = 0.0;
      st.ent_u.reset();
      double tr_term = 0.0;
      if (p.phases & I2C_PH_FORWARD) {
        __threadfence_block();
        gsync();
        prefetch_rec(wk.rec(wk.prior, 0, LY::E_POST), LY::P_KK, 0);
        if (!load_x0()) wk.fail(I2C_FAIL_CHOL_PRIOR, it, 0);
        for (int t = 0; t < T; ++t) {
          I2C_TICK(14)
          rec_wait();
          I2C_TICK(13)
          if (t + 1 < T) prefetch_rec(wk.rec(wk.prior, t + 1, LY::E_POST), LY::P_KK, (t + 1) & 1);
          forward_cell(it, t, wk.staged_flags(nullptr, t, flipped), alpha, aux, rec_buf(t), ent_x);
        }
      }
      if (p.phases & I2C_PH_BACKWARD) {
        __threadfence_block();
        gsync();
        prefetch_rec(wk.rec(p.filt, T - 1, LY::E_FILT), LY::E_FILT, (T - 1) & 1);
        if (!(p.phases & I2C_PH_FORWARD)) {
          // resume from the stored filtered message of the last cell
          const double* fr = wk.rec(p.filt, T - 1, LY::E_FILT);
          load_carry(fr + (size_t)LY::F_MU3 * TILE, fr + (size_t)LY::F_SIG3 * TILE);
        }
        backward_terminal(it, T - 1, temp, wk.cell_alpha(T - 1, p.cell_flags[wk.slot(T - 1)], alpha), tr_term);
        for (int t = T - 1; t >= 0; --t) {
          I2C_TICK(14)
          rec_wait();
          I2C_TICK(13)
          if (t > 0) prefetch_rec(wk.rec(p.filt, t - 1, LY::E_FILT), LY::E_FILT, (t - 1) & 1);
          backward_cell(it, t, aux, rec_buf(t), st);
        }
        if (p.cov_ctrl) temp += p.dtemp;
        wk.latest = wk.post;
      }
      PStats ps;
      ps.cost = ps.cost_var = ps.tr = 0.0;
      ps.cost_min = INFINITY;
      ps.ent.reset();
      if (p.phases & I2C_PH_PROPAGATE) {
        __threadfence_block();  // posterior records written by other lanes of the group are read back
        gsync();
        prefetch_rec(wk.rec(wk.latest, 0, LY::E_POST), LY::P_KK, 0);
        if (!load_x0()) wk.fail(I2C_FAIL_CHOL_PROPAGATE, it, 0);
        for (int t = 0; t < T; ++t) {
          rec_wait();
          if (t + 1 < T) prefetch_rec(wk.rec(wk.latest, t + 1, LY::E_POST), LY::P_KK, (t + 1) & 1);
          propagate_cell(it, t, wk.staged_flags(nullptr, t, flipped), aux, rec_buf(t), ps);
        }
        if (p.cov_ctrl && writer) {
          // KL(N(mu_x3_pf, sig_x3_pf) || N(mu_xT, sig_xT)) of the last cell (i2c.py:1012-1019, 1223-1229)
          const double* L0 = sm + GL::O_L0;
          double A[TRI(DX)], ia[DX], d[DX];
#pragma unroll
          for (int i = 0; i < TRI(DX); ++i) A[i] = p.sxt[i];
          chol_rows<DX>(A, ia);
#pragma unroll
          for (int i = 0; i < DX; ++i) d[i] = p.mu_xt[i] - sm[GL::O_M0 + i];
          fwd_subst<DX>(A, ia, d);
          double dist = 0.0, tr = 0.0, ld1 = 0.0;
#pragma unroll
          for (int i = 0; i < DX; ++i) dist = fma(d[i], d[i], dist);
#pragma unroll
          for (int j = 0; j < DX; ++j) {
            double w[DX];
#pragma unroll
            for (int i = 0; i < DX; ++i) w[i] = (i >= j) ? L0[i * DX + j] : 0.0;
            fwd_subst<DX>(A, ia, w);
#pragma unroll
            for (int i = 0; i < DX; ++i) tr = fma(w[i], w[i], tr);
            ld1 += log(L0[j * DX + j]);
          }
          wk.metric(I2C_M_KL_TERM, it, 0.5 * (p.sxt_logdet - 2.0 * ld1 + tr + dist - (double)DX));
        }
      }
      if ((p.phases & I2C_PH_MSTEP) && writer) {
        wk.metric(I2C_M_COST_M, it, st.cost);
        wk.metric(I2C_M_COST_M_VAR, it, st.cost_var);
        if (p.phases & I2C_PH_PROPAGATE) {
          wk.metric(I2C_M_COST_PF, it, ps.cost);
          wk.metric(I2C_M_COST_PF_VAR, it, ps.cost_var);
          wk.metric(I2C_M_COST_PF_MIN, it, ps.cost_min);
          wk.metric(I2C_M_ALPHA_PF, it, ps.tr / (double)(DZ * T));
          wk.metric(I2C_M_PROPAGATE_ENTROPY, it, (double)(T * DX) * HALF_LOG_2PIE + ps.ent.value());
        } else {
          wk.metric(I2C_M_COST_PF, it, -1.0);
        }
        wk.metric(I2C_M_POLICY_ENTROPY, it, (double)(T * DU) * HALF_LOG_2PIE + st.ent_u.value());
        wk.metric(I2C_M_X_PRIOR_ENTROPY, it, (double)(T * DX) * HALF_LOG_2PIE + ent_x.value());
      }
      if (p.phases & I2C_PH_UPDATE_PRIORS) {
        if (wk.latest == wk.post) {
          double* tmp = wk.prior;
          wk.prior = wk.post;
          wk.post = tmp;
        }
        flipped = true;
      }
      if (p.phases & I2C_PH_MSTEP) {
        alpha = wk.mstep_alpha(it, st.tr, tr_term, alpha);  // identical in every lane (metrics stored redundantly)
        wk.own_alpha_valid = false;
      }
      if (p.phases & I2C_PH_CALIBRATE) {
        const double a_pf = ps.tr / (double)(DZ * T);
        const bool upd = (p.phases & I2C_PH_ONLY_DECREASE) ? (a_pf < alpha) : true;
        if (upd) {
          alpha = a_pf;
          wk.own_alpha_valid = false;
        }
        if (writer) wk.metric(I2C_M_ALPHA, it, alpha);
      }
    }
#ifdef I2C_GROUP_TIMING
    if (wk.b == 0 && r == 0) {
      const double cells = (double)p.n_iter * T;
      printf("group timing (cycles / cell): joint %.0f obsT %.0f cond %.0f cholF %.0f dynT %.0f cholX+J %.0f term+store %.0f | head %.0f cholP %.0f K %.0f | obsT %.0f stats %.0f trace %.0f | recwait %.0f other %.0f\n",
             tacc[0] / cells, tacc[1] / cells, tacc[2] / cells, tacc[3] / cells, tacc[4] / cells, tacc[5] / cells, tacc[6] / cells,
             tacc[7] / cells, tacc[8] / cells, tacc[9] / cells, tacc[12] / cells, tacc[10] / cells, tacc[11] / cells, tacc[13] / cells, tacc[14] / cells);
    }
#endif
    if (!writer) return;  // the checks are group-uniform: every lane saw the same status
    p.alpha[wk.b] = alpha;
    if (wk.status != I2C_OK && p.status[wk.b] == I2C_OK) {
      p.status[wk.b] = wk.status;
      p.info[wk.b] = wk.info;
    }
  }
#undef I2C_FOR_ROWS
};

template <class Env, int G, int WPB>
__global__ void __launch_bounds__(WPB * TILE) em_group_kernel(const __grid_constant__ KParams pin) {
  constexpr int PPW = TILE / G;  // problems per warp
  const int warp = blockIdx.x * WPB + threadIdx.x / TILE, lane = threadIdx.x % TILE;
  const int nwarps = pin.ntiles * G;
  if (warp >= nwarps) return;
  extern __shared__ __align__(16) double group_smem[];
  const int tile = warp / G, sub = warp % G;
  const int pl = sub * PPW + lane / G;  // problem slot inside the tile
  double* sm = group_smem + ((size_t)(threadIdx.x / TILE) * PPW + lane / G) * GroupLay<Env>::SIZE;
  GroupWorker<Env, G> gw(pin, tile, pl, lane % G, sm);
  gw.run();
}

template <class Env, int G>
static int launch_em_group(const KParams& p, cudaStream_t s) {
  constexpr int WPB = 2;
  constexpr int PPW = TILE / G;
  const size_t smem = (size_t)WPB * PPW * GroupLay<Env>::SIZE * sizeof(double);
  if (int e = allow_big_smem<em_group_kernel<Env, G, WPB>>()) return e;
  KParams q = p;
  q.stage_meta = 0;
  const int nwarps = p.ntiles * G;
  em_group_kernel<Env, G, WPB><<<(nwarps + WPB - 1) / WPB, WPB * TILE, smem, s>>>(q);
  return (int)cudaGetLastError();
}

// Policy (measured, profiles/r01h_group_vs_thread.txt): the group kernel wins where the per-thread kernel is issue-bound on
// one sub-partition per tile AND heavy per problem -- the double cart-pole (n = 7: 5.7 KB of register spills per thread,
// ~350 instructions per sigma-point evaluation): 1.54x at 512 / 2048 problems, 1.33x at 4096, slower from 8192 on.  For the
// quadrotor, cart-pole and pendulum the per-thread kernels are 1.5-3x faster at every batch size (cheap dynamics, no
// spills), so the variant is taken automatically only for n = 7 up to 160 tiles.  I2C_B200_GROUP=0/1 forces the choice,
// I2C_B200_GROUP_MAX_TILES overrides the threshold.
template <class Env>
static int launch_em_group_maybe(const KParams& p, cudaStream_t s) {
  constexpr int N = Lay<Env>::N;
  constexpr int G = N >= 5 ? 8 : 4;
  if (p.linearize || p.gh.degree > 0 || (p.phases & I2C_PH_RICCATI)) return kGroupNotTaken;
  if (p.group_mode == 0) return kGroupNotTaken;
  const int max_tiles = p.group_max_tiles > 0 ? p.group_max_tiles : (N == 7 ? 160 : 0);
  if (p.group_mode < 0 && p.ntiles > max_tiles) return kGroupNotTaken;
  // round 2: with the HOT specialisation (branch-free + angle-addition trig: 11 instead of 34 sincos per double cart-pole
  // dynamics transform, no per-cell flag branches) the per-thread team kernel overtook this one on the config-4 shard:
  // 6.9 ms against 10.6 ms per EM iteration at 2048 x 500 (profiles/r02_env_sweep.jsonl).  Auto mode keeps the group kernel
  // for the configurations the HOT kernels do not cover (auxiliary records, per-problem targets, per-cell alpha).
  if (p.group_mode < 0 && p.hot) return kGroupNotTaken;
  return launch_em_group<Env, G>(p, s);
}

}  // namespace i2c
