// Parallel-in-time variant of the EM sweep (north-star "parallel-in-time associative-scan formulation of the filter
// and smoother for long horizons").  NOT in the reference: its sweeps are strictly
// sequential Python loops (i2c/i2c.py:876-886).  Exact (same Gaussians up to round-off) whenever the cell maps are
// linear-Gaussian in the state message: Linearize inference on the linear environments with cells that are either
// state-action independent or use the un-weighted feedback prior (no expert pdf ratio).
//
// Formulation.  Every cell t maps the incoming state message N(m, S) to the outgoing one.  For the cells above that map is
// the Kalman "filtering element" of Sarkka & Garcia-Fernandez (2021), a = (A, b, C, eta, J):
//     p(x_{t+1} | x_t, z_t) = N(A x_t + b, C),     p(z_t | x_t)  ~  N_info(x_t; eta, J)
// and the smoother step is the affine element (E, g, L):  m_t = E m_{t+1} + g,  S_t = E S_{t+1} E^T + L.  Both compose
// associatively, so the horizon is cut into chunks that are processed concurrently:
//     1. *_local   one warp per (tile, chunk): build the elements of the chunk's cells and compose them   (parallel)
//     2. *_prefix  one warp per tile: push the boundary message through the chunk aggregates              (n_chunks steps)
//     3. *_cells   one warp per (tile, chunk): the ordinary sequential cell code (Worker::forward_cell /
//                  backward_cell) started from the chunk's exact boundary message                         (parallel)
// Step 3 writes exactly the records of the sequential kernel, so every getter, the Riccati sweep and propagate work
// unchanged afterwards.  Sequential depth: 2 * chunk + n_chunks cells instead of T.
#pragma once
#include "i2c_kernels.cuh"

namespace i2c {

template <int DX>
struct ScanDims {
  static constexpr int X2 = DX * DX;
  static constexpr int E_FAGG = 3 * X2 + 2 * DX, E_BAGG = 2 * X2 + DX, E_MSG = DX + TRI(DX);
};
__host__ __device__ constexpr int scan_e_fagg(int dx) { return 3 * dx * dx + 2 * dx; }
__host__ __device__ constexpr int scan_e_bagg(int dx) { return 2 * dx * dx + dx; }
__host__ __device__ constexpr int scan_e_msg(int dx) { return dx + dx * (dx + 1) / 2; }

// filtering element; C and J kept as full row-major symmetric matrices
template <int DX>
struct FElem {
  double A[DX * DX], b[DX], C[DX * DX], eta[DX], J[DX * DX];
};
template <int DX>
struct BElem {
  double E[DX * DX], g[DX], L[DX * DX];
};

template <int DX>
__device__ __forceinline__ void symmetrise(double* M) {
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) {
      const double v = 0.5 * (M[i * DX + j] + M[j * DX + i]);
      M[i * DX + j] = v;
      M[j * DX + i] = v;
    }
}

// a_i (earlier) followed by a_j (later)  ->  out   (Sarkka & Garcia-Fernandez 2021, Lemma 8)
template <int DX>
__device__ __forceinline__ void combine_f(const FElem<DX>& ai, const FElem<DX>& aj, FElem<DX>& out) {
  constexpr int X2 = DX * DX;
  double M[X2];  // (I + C_i J_j)^-1
  mm<DX, DX, DX>(ai.C, aj.J, M);
#pragma unroll
  for (int i = 0; i < DX; ++i) M[i * DX + i] += 1.0;
  inv_gj<DX>(M);
  double AjM[X2], T1[X2], v[DX], w[DX];
  mm<DX, DX, DX>(aj.A, M, AjM);
  // b_ij = A_j M (b_i + C_i eta_j) + b_j
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = ai.b[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(ai.C[i * DX + k], aj.eta[k], s);
    v[i] = s;
  }
  FElem<DX> o;
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = aj.b[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(AjM[i * DX + k], v[k], s);
    o.b[i] = s;
  }
  mm<DX, DX, DX>(AjM, ai.A, o.A);
  // C_ij = A_j M C_i A_j^T + C_j
  mm<DX, DX, DX>(AjM, ai.C, T1);
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = aj.C[i * DX + j];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(T1[i * DX + k], aj.A[j * DX + k], s);
      o.C[i * DX + j] = s;
    }
  symmetrise<DX>(o.C);
  // eta_ij = A_i^T M^T (eta_j - J_j b_i) + eta_i ;  J_ij = A_i^T M^T J_j A_i + J_i      ((I + J_j C_i)^-1 = M^T)
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = aj.eta[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(-aj.J[i * DX + k], ai.b[k], s);
    v[i] = s;
  }
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(M[k * DX + i], v[k], s);
    w[i] = s;
  }
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = ai.eta[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(ai.A[k * DX + i], w[k], s);
    o.eta[i] = s;
  }
  double MtJ[X2], T2[X2];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(M[k * DX + i], aj.J[k * DX + j], s);
      MtJ[i * DX + j] = s;
    }
  mm<DX, DX, DX>(MtJ, ai.A, T2);
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = ai.J[i * DX + j];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(ai.A[k * DX + i], T2[k * DX + j], s);
      o.J[i * DX + j] = s;
    }
  symmetrise<DX>(o.J);
  out = o;
}

// push the message N(m, S) through a filtering element (= combine with the element (0, m, S, 0, 0))
template <int DX>
__device__ __forceinline__ void apply_f(const FElem<DX>& a, double* m, double* S /* packed lower */) {
  constexpr int X2 = DX * DX;
  double Sf[X2], M[X2], v[DX], AM[X2], T1[X2];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) Sf[i * DX + j] = S[six(i, j)];
  mm<DX, DX, DX>(Sf, a.J, M);
#pragma unroll
  for (int i = 0; i < DX; ++i) M[i * DX + i] += 1.0;
  inv_gj<DX>(M);
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = m[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(Sf[i * DX + k], a.eta[k], s);
    v[i] = s;
  }
  mm<DX, DX, DX>(a.A, M, AM);
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = a.b[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(AM[i * DX + k], v[k], s);
    m[i] = s;
  }
  mm<DX, DX, DX>(AM, Sf, T1);
  double So[X2];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = a.C[i * DX + j];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(T1[i * DX + k], a.A[j * DX + k], s);
      So[i * DX + j] = s;
    }
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) S[tix(i, j)] = 0.5 * (So[i * DX + j] + So[j * DX + i]);
}

// smoother elements: `l` is the cell to the LEFT (applied last), `r` the aggregate of the cells to its right
template <int DX>
__device__ __forceinline__ void combine_b(const BElem<DX>& l, const BElem<DX>& r, BElem<DX>& out) {
  constexpr int X2 = DX * DX;
  BElem<DX> o;
  double T1[X2];
  mm<DX, DX, DX>(l.E, r.E, o.E);
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = l.g[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(l.E[i * DX + k], r.g[k], s);
    o.g[i] = s;
  }
  mm<DX, DX, DX>(l.E, r.L, T1);
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = l.L[i * DX + j];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(T1[i * DX + k], l.E[j * DX + k], s);
      o.L[i * DX + j] = s;
    }
  symmetrise<DX>(o.L);
  out = o;
}
template <int DX>
__device__ __forceinline__ void apply_b(const BElem<DX>& a, double* m, double* S /* packed lower */) {
  constexpr int X2 = DX * DX;
  double Sf[X2], T1[X2], mo[DX];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) Sf[i * DX + j] = S[six(i, j)];
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    double s = a.g[i];
#pragma unroll
    for (int k = 0; k < DX; ++k) s = fma(a.E[i * DX + k], m[k], s);
    mo[i] = s;
  }
  mm<DX, DX, DX>(a.E, Sf, T1);
  double So[X2];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      double s = a.L[i * DX + j];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(T1[i * DX + k], a.E[j * DX + k], s);
      So[i * DX + j] = s;
    }
#pragma unroll
  for (int i = 0; i < DX; ++i) m[i] = mo[i];
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) S[tix(i, j)] = 0.5 * (So[i * DX + j] + So[j * DX + i]);
}

template <class Env>
struct ScanWorker {
  using WK = Worker<Env, true, true>;
  using LY = Lay<Env>;
  static constexpr int DX = LY::DX, DU = LY::DU, N = LY::N, DZ = LY::DZ, DZT = LY::DZT, X2 = DX * DX;
  using SD = ScanDims<DX>;

  // The filtering element of cell t: the exact linear-Gaussian restatement of Worker::forward_cell in Linearize mode on
  // a linear environment (_forward_msgs_linearize, i2c.py:244-348, with K un-weighted).  With s = (x, u):
  //   s = G x + g + w,  G = [I; K],  g = [0; k~],  Cov(w) = blockdiag(0, S_u~)        (prior / feedback prior)
  //   z = H s + xi,  Cov(xi) = alpha QR^-1                                           (cost observation)
  //   x' = F s + a + eta,  F = [A B]                                                 (dynamics)
  __device__ static void forward_elem(const WK& wk, int t, int flags, double alpha, FElem<DX>& e) {
    const KParams& p = wk.p;
    const double* pr = wk.rec(wk.prior, t, LY::E_POST);
    double ktil[DU], Sut[DU * DU], Kt[DU * DX];
#pragma unroll
    for (int r = 0; r < DU; ++r) ktil[r] = pr[(LY::P_MU + DX + r) * TILE];
#pragma unroll
    for (int r = 0; r < DU; ++r)
#pragma unroll
      for (int q = 0; q < DU; ++q) Sut[r * DU + q] = pr[(LY::P_SIG + six(DX + r, DX + q)) * TILE];
#pragma unroll
    for (int i = 0; i < DU * DX; ++i) Kt[i] = 0.0;
    if (!(flags & I2C_CELL_INDEPENDENT)) {
      // u | x = mu_u0_m + K (x - mu_x0_m) + w,  Cov(w) = sig_u0_m - K sig_ux^T   (i2c.py:266-276, un-weighted K)
      double mx[DX];
#pragma unroll
      for (int i = 0; i < DX; ++i) mx[i] = pr[(LY::P_MU + i) * TILE];
#pragma unroll
      for (int i = 0; i < DU * DX; ++i) Kt[i] = pr[(LY::P_K + i) * TILE];
#pragma unroll
      for (int r = 0; r < DU; ++r) {
#pragma unroll
        for (int k = 0; k < DX; ++k) ktil[r] = fma(-Kt[r * DX + k], mx[k], ktil[r]);
#pragma unroll
        for (int q = 0; q <= r; ++q) {
          double v = Sut[r * DU + q];
#pragma unroll
          for (int k = 0; k < DX; ++k) v = fma(-Kt[r * DX + k], pr[(LY::P_SIG + tix(DX + q, k)) * TILE], v);
          Sut[r * DU + q] = v;
          Sut[q * DU + r] = v;
        }
      }
    }
    // dynamics parameters (EnvLinear layout: A row-major, B, a)
    double A[X2], Bm[DX * DU], av[DX];
#pragma unroll
    for (int i = 0; i < X2; ++i) A[i] = wk.par[i];
#pragma unroll
    for (int i = 0; i < DX * DU; ++i) Bm[i] = wk.par[X2 + i];
#pragma unroll
    for (int i = 0; i < DX; ++i) av[i] = wk.par[X2 + DX * DU + i];
    // observation: z_a = s[src(a)]
    const double a_cell = wk.cell_alpha(t, flags, alpha);
    double HG[DZ * DX], Hg[DZ], Rz[TRI(DZ)], Cxz[DX * DZ], z[DZ];
#pragma unroll
    for (int a = 0; a < DZ; ++a) {
      const int sa = Env::obs_src(a);
#pragma unroll
      for (int j = 0; j < DX; ++j) HG[a * DX + j] = sa < DX ? (sa == j ? 1.0 : 0.0) : Kt[(sa >= DX ? sa - DX : 0) * DX + j];
      Hg[a] = sa < DX ? 0.0 : ktil[sa >= DX ? sa - DX : 0];
#pragma unroll
      for (int bb = 0; bb <= a; ++bb) {
        const int sb = Env::obs_src(bb);
        double v = a_cell * p.QRinv[a * DZ + bb];
        if (sa >= DX && sb >= DX) v += Sut[(sa >= DX ? sa - DX : 0) * DU + (sb >= DX ? sb - DX : 0)];
        Rz[tix(a, bb)] = v;
      }
#pragma unroll
      for (int i = 0; i < DX; ++i) {
        double v = 0.0;
        if (sa >= DX) {
#pragma unroll
          for (int k = 0; k < DU; ++k) v = fma(Bm[i * DU + k], Sut[k * DU + (sa >= DX ? sa - DX : 0)], v);
        }
        Cxz[i * DZ + a] = v;
      }
    }
    wk.load_z(t, z);
    double invd[DZ];
    chol_rows<DZ>(Rz, invd);  // alpha QR^-1 is positive definite; a failure shows up in the cell pass of stage 3
    double r[DZ], P[DX * DZ], Q[DX * DZ];  // P[j][:] = Lz^-1 HG[:, j],  Q[i][:] = Lz^-1 Cxz[i, :]
#pragma unroll
    for (int a = 0; a < DZ; ++a) r[a] = z[a] - Hg[a];
    fwd_subst<DZ>(Rz, invd, r);
#pragma unroll
    for (int j = 0; j < DX; ++j) {
#pragma unroll
      for (int a = 0; a < DZ; ++a) {
        P[j * DZ + a] = HG[a * DX + j];
        Q[j * DZ + a] = Cxz[j * DZ + a];
      }
      fwd_subst<DZ>(Rz, invd, P + j * DZ);
      fwd_subst<DZ>(Rz, invd, Q + j * DZ);
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      double se = 0.0, sb = av[i];
#pragma unroll
      for (int a = 0; a < DZ; ++a) {
        se = fma(P[i * DZ + a], r[a], se);
        sb = fma(Q[i * DZ + a], r[a], sb);
      }
#pragma unroll
      for (int k = 0; k < DU; ++k) sb = fma(Bm[i * DU + k], ktil[k], sb);
      e.eta[i] = se;
      e.b[i] = sb;
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        double sj = 0.0, sa_ = A[i * DX + j], sc = p.sig_eta[six(i, j)];
#pragma unroll
        for (int a = 0; a < DZ; ++a) {
          sj = fma(P[i * DZ + a], P[j * DZ + a], sj);
          sa_ = fma(-Q[i * DZ + a], P[j * DZ + a], sa_);
          sc = fma(-Q[i * DZ + a], Q[j * DZ + a], sc);
        }
#pragma unroll
        for (int k = 0; k < DU; ++k) {
          sa_ = fma(Bm[i * DU + k], Kt[k * DX + j], sa_);
#pragma unroll
          for (int l = 0; l < DU; ++l) sc = fma(Bm[i * DU + k] * Sut[k * DU + l], Bm[j * DU + l], sc);
        }
        e.J[i * DX + j] = sj;
        e.A[i * DX + j] = sa_;
        e.C[i * DX + j] = sc;
      }
    }
  }

  // RTS element of cell t from its filtered record (Worker::backward_head restricted to the state block)
  __device__ static void backward_elem(const WK& wk, int t, BElem<DX>& e) {
    const double* fr = wk.rec(wk.p.filt, t, LY::E_FILT);
    double m3[DX], S3[X2], T1[X2];
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      m3[i] = fr[(LY::F_MU3 + i) * TILE];
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        e.E[i * DX + j] = fr[(LY::F_J + i * DX + j) * TILE];
        S3[i * DX + j] = fr[(LY::F_SIG3 + six(i, j)) * TILE];
      }
    }
    mm<DX, DX, DX>(e.E, S3, T1);
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      double s = fr[(LY::F_MU1 + i) * TILE];
#pragma unroll
      for (int k = 0; k < DX; ++k) s = fma(-e.E[i * DX + k], m3[k], s);
      e.g[i] = s;
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        double v = fr[(LY::F_SIG1 + six(i, j)) * TILE];
#pragma unroll
        for (int k = 0; k < DX; ++k) v = fma(-T1[i * DX + k], e.E[j * DX + k], v);
        e.L[i * DX + j] = v;
      }
    }
  }

  __device__ static double* slotp(double* base, int E, int ch, int tile, int lane, int ntiles) {
    return base + ((size_t)ch * ntiles + tile) * (size_t)E * TILE + lane;
  }
  __device__ static void merge_status(const WK& wk) {
    if (wk.status != I2C_OK && atomicCAS(&wk.p.status[wk.b], (int)I2C_OK, wk.status) == I2C_OK) wk.p.info[wk.b] = wk.info;
  }
};

template <class Env, int STAGE>
__global__ void __launch_bounds__(64) scan_kernel(const __grid_constant__ KParams p, const __grid_constant__ ScanArgs a) {
  using SW = ScanWorker<Env>;
  using WK = typename SW::WK;
  using LY = Lay<Env>;
  constexpr int DX = LY::DX, N = LY::N, X2 = DX * DX;
  using SD = ScanDims<DX>;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) / TILE, lane = threadIdx.x % TILE;
  constexpr bool PER_CHUNK = STAGE == SCAN_FWD_LOCAL || STAGE == SCAN_FWD_CELLS || STAGE == SCAN_BWD_LOCAL || STAGE == SCAN_BWD_CELLS;
  const int nw = PER_CHUNK ? p.ntiles * a.n_chunks : p.ntiles;
  if (gw >= nw) return;
  const int tile = PER_CHUNK ? gw % p.ntiles : gw, ch = PER_CHUNK ? gw / p.ntiles : 0;
  WK wk(p, tile, lane, nullptr, nullptr);
  const int t0 = ch * a.chunk, t1 = min(p.T, t0 + a.chunk);
  const bool aux = p.phases & I2C_PH_STORE_AUX;
  const double alpha = p.alpha[wk.b];

  if constexpr (STAGE == SCAN_FWD_LOCAL) {
    if (ch == a.n_chunks - 1) return;  // nothing lies to the right of the last chunk
    FElem<DX> agg;
    SW::forward_elem(wk, t0, wk.staged_flags(nullptr, t0, false), alpha, agg);
    for (int t = t0 + 1; t < t1; ++t) {
      FElem<DX> e;
      SW::forward_elem(wk, t, wk.staged_flags(nullptr, t, false), alpha, e);
      combine_f<DX>(agg, e, agg);
    }
    double* o = SW::slotp(a.fagg, SD::E_FAGG, ch, tile, lane, p.ntiles);
#pragma unroll
    for (int i = 0; i < X2; ++i) {
      o[i * TILE] = agg.A[i];
      o[(X2 + i) * TILE] = agg.C[i];
      o[(2 * X2 + i) * TILE] = agg.J[i];
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      o[(3 * X2 + i) * TILE] = agg.b[i];
      o[(3 * X2 + DX + i) * TILE] = agg.eta[i];
    }
  } else if constexpr (STAGE == SCAN_FWD_PREFIX) {
    Carry<DX> c;
    wk.load_x0(c);
    for (int k = 0; k < a.n_chunks; ++k) {
      double* o = SW::slotp(a.cin, SD::E_MSG, k, tile, lane, p.ntiles);
#pragma unroll
      for (int i = 0; i < DX; ++i) o[i * TILE] = c.m[i];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) o[(DX + i) * TILE] = c.S[i];
      if (k + 1 < a.n_chunks) {
        const double* q = SW::slotp(a.fagg, SD::E_FAGG, k, tile, lane, p.ntiles);
        FElem<DX> e;
#pragma unroll
        for (int i = 0; i < X2; ++i) {
          e.A[i] = q[i * TILE];
          e.C[i] = q[(X2 + i) * TILE];
          e.J[i] = q[(2 * X2 + i) * TILE];
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) {
          e.b[i] = q[(3 * X2 + i) * TILE];
          e.eta[i] = q[(3 * X2 + DX + i) * TILE];
        }
        apply_f<DX>(e, c.m, c.S);
      }
    }
  } else if constexpr (STAGE == SCAN_FWD_CELLS) {
    Carry<DX> c;
    const double* q = SW::slotp(a.cin, SD::E_MSG, ch, tile, lane, p.ntiles);
#pragma unroll
    for (int i = 0; i < DX; ++i) c.m[i] = q[i * TILE];
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) {
      c.S[i] = q[(DX + i) * TILE];
      c.L[i] = c.S[i];
    }
    if (!chol_rows<DX>(c.L, c.invd)) wk.fail(I2C_FAIL_CHOL_PRIOR, a.it, t0);
    LogAcc ent_x;
    ent_x.reset();
    for (int t = t0; t < t1; ++t)
      wk.forward_cell(a.it, t, wk.staged_flags(nullptr, t, false), alpha, aux, wk.rec(wk.prior, t, LY::E_POST), c, ent_x);
    double* pt = a.part + (size_t)ch * SCAN_PARTS * p.Bpad + wk.b;
    pt[(size_t)SP_ENTX_M * p.Bpad] = ent_x.m;
    pt[(size_t)SP_ENTX_E * p.Bpad] = (double)ent_x.e;
    SW::merge_status(wk);
  } else if constexpr (STAGE == SCAN_BWD_LOCAL) {
    if (ch == 0) return;  // nothing lies to the left of the first chunk
    BElem<DX> agg;
    SW::backward_elem(wk, t1 - 1, agg);
    for (int t = t1 - 2; t >= t0; --t) {
      BElem<DX> e;
      SW::backward_elem(wk, t, e);
      combine_b<DX>(e, agg, agg);
    }
    double* o = SW::slotp(a.bagg, SD::E_BAGG, ch, tile, lane, p.ntiles);
#pragma unroll
    for (int i = 0; i < X2; ++i) {
      o[i * TILE] = agg.E[i];
      o[(X2 + i) * TILE] = agg.L[i];
    }
#pragma unroll
    for (int i = 0; i < DX; ++i) o[(2 * X2 + i) * TILE] = agg.g[i];
  } else if constexpr (STAGE == SCAN_BWD_SUFFIX) {
    const int T = p.T;
    Carry<DX> c;
    const double* fr = wk.rec(p.filt, T - 1, LY::E_FILT);
#pragma unroll
    for (int i = 0; i < DX; ++i) c.m[i] = fr[(LY::F_MU3 + i) * TILE];
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) {
      c.S[i] = fr[(LY::F_SIG3 + i) * TILE];
      c.L[i] = c.S[i];
    }
    chol_rows<DX>(c.L, c.invd);
    double m3m[DX], S3m[TRI(DX)], tr_term = 0.0;
    wk.backward_terminal(a.it, T - 1, a.temp, wk.cell_alpha(T - 1, p.cell_flags[wk.slot(T - 1)], alpha), c, m3m, S3m, tr_term);
    a.tail[wk.b] = tr_term;
    for (int k = a.n_chunks - 1; k >= 0; --k) {
      double* o = SW::slotp(a.bin, SD::E_MSG, k, tile, lane, p.ntiles);
#pragma unroll
      for (int i = 0; i < DX; ++i) o[i * TILE] = m3m[i];
#pragma unroll
      for (int i = 0; i < TRI(DX); ++i) o[(DX + i) * TILE] = S3m[i];
      if (k > 0) {
        const double* q = SW::slotp(a.bagg, SD::E_BAGG, k, tile, lane, p.ntiles);
        BElem<DX> e;
#pragma unroll
        for (int i = 0; i < X2; ++i) {
          e.E[i] = q[i * TILE];
          e.L[i] = q[(X2 + i) * TILE];
        }
#pragma unroll
        for (int i = 0; i < DX; ++i) e.g[i] = q[(2 * X2 + i) * TILE];
        apply_b<DX>(e, m3m, S3m);
      }
    }
    SW::merge_status(wk);
  } else if constexpr (STAGE == SCAN_BWD_CELLS) {
    double m3m[DX], S3m[TRI(DX)];
    const double* q = SW::slotp(a.bin, SD::E_MSG, ch, tile, lane, p.ntiles);
#pragma unroll
    for (int i = 0; i < DX; ++i) m3m[i] = q[i * TILE];
#pragma unroll
    for (int i = 0; i < TRI(DX); ++i) S3m[i] = q[(DX + i) * TILE];
    typename WK::Stats st;
    st.cost = st.cost_var = st.tr = 0.0;
    st.ent_u.reset();
    for (int t = t1 - 1; t >= t0; --t)
      wk.backward_cell(a.it, t, wk.staged_flags(nullptr, t, false), aux, wk.rec(p.filt, t, LY::E_FILT), m3m, S3m, st);
    double* pt = a.part + (size_t)ch * SCAN_PARTS * p.Bpad + wk.b;
    pt[(size_t)SP_COST * p.Bpad] = st.cost;
    pt[(size_t)SP_COST_VAR * p.Bpad] = st.cost_var;
    pt[(size_t)SP_TR * p.Bpad] = st.tr;
    pt[(size_t)SP_ENTU_M * p.Bpad] = st.ent_u.m;
    pt[(size_t)SP_ENTU_E * p.Bpad] = (double)st.ent_u.e;
    SW::merge_status(wk);
  } else {  // SCAN_MSTEP: fixed-order reduction of the chunk statistics, metrics, alpha update (Worker::run_impl tail)
    const double HALF_LOG_2PIE = 1.4189385332046727;
    typename WK::Stats st;
    st.cost = st.cost_var = st.tr = 0.0;
    st.ent_u.reset();
    LogAcc ent_x;
    ent_x.reset();
    for (int k = 0; k < a.n_chunks; ++k) {
      const double* pt = a.part + (size_t)k * SCAN_PARTS * p.Bpad + wk.b;
      ent_x.mul(pt[(size_t)SP_ENTX_M * p.Bpad]);
      ent_x.e += (int)pt[(size_t)SP_ENTX_E * p.Bpad];
    }
    for (int k = a.n_chunks - 1; k >= 0; --k) {
      const double* pt = a.part + (size_t)k * SCAN_PARTS * p.Bpad + wk.b;
      st.cost += pt[(size_t)SP_COST * p.Bpad];
      st.cost_var += pt[(size_t)SP_COST_VAR * p.Bpad];
      st.tr += pt[(size_t)SP_TR * p.Bpad];
      st.ent_u.mul(pt[(size_t)SP_ENTU_M * p.Bpad]);
      st.ent_u.e += (int)pt[(size_t)SP_ENTU_E * p.Bpad];
    }
    wk.metric(I2C_M_COST_M, a.it, st.cost);
    wk.metric(I2C_M_COST_M_VAR, a.it, st.cost_var);
    wk.metric(I2C_M_COST_PF, a.it, -1.0);
    wk.metric(I2C_M_POLICY_ENTROPY, a.it, (double)(p.T * LY::DU) * HALF_LOG_2PIE + st.ent_u.value());
    wk.metric(I2C_M_X_PRIOR_ENTROPY, a.it, (double)(p.T * DX) * HALF_LOG_2PIE + ent_x.value());
    p.alpha[wk.b] = wk.mstep_alpha(a.it, st.tr, a.tail[wk.b], alpha);
    SW::merge_status(wk);
  }
}

template <class Env>
static int launch_scan_t(int stage, const KParams& p, const ScanArgs& a, cudaStream_t s) {
  if constexpr (Env::LINEAR) {
    KParams q = p;
    q.stage_meta = 0;
    const int threads = 64;
    const long long per_chunk = (long long)p.ntiles * a.n_chunks * TILE, per_tile = (long long)p.ntiles * TILE;
    auto blocks = [&](long long n) { return (unsigned)((n + threads - 1) / threads); };
    switch (stage) {
      case SCAN_FWD_LOCAL: scan_kernel<Env, SCAN_FWD_LOCAL><<<blocks(per_chunk), threads, 0, s>>>(q, a); break;
      case SCAN_FWD_PREFIX: scan_kernel<Env, SCAN_FWD_PREFIX><<<blocks(per_tile), threads, 0, s>>>(q, a); break;
      case SCAN_FWD_CELLS: scan_kernel<Env, SCAN_FWD_CELLS><<<blocks(per_chunk), threads, 0, s>>>(q, a); break;
      case SCAN_BWD_LOCAL: scan_kernel<Env, SCAN_BWD_LOCAL><<<blocks(per_chunk), threads, 0, s>>>(q, a); break;
      case SCAN_BWD_SUFFIX: scan_kernel<Env, SCAN_BWD_SUFFIX><<<blocks(per_tile), threads, 0, s>>>(q, a); break;
      case SCAN_BWD_CELLS: scan_kernel<Env, SCAN_BWD_CELLS><<<blocks(per_chunk), threads, 0, s>>>(q, a); break;
      case SCAN_MSTEP: scan_kernel<Env, SCAN_MSTEP><<<blocks(per_tile), threads, 0, s>>>(q, a); break;
      default: return -1;
    }
    return (int)cudaGetLastError();
  } else {
    return -2;  // the cell maps of the nonlinear environments are not linear-Gaussian: no exact scan exists
  }
}

}  // namespace i2c
