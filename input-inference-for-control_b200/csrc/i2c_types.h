// Internal types shared by the kernels and the C-ABI implementation (not part of the public header).
#pragma once
#include <stdint.h>

#include "../../include/i2c_b200.h"

namespace i2c {

constexpr int TILE = 32;       // problems per tile == one warp; innermost (contiguous) axis of every record
// Gauss-Hermite rule (exp_types.py:52-68): 1-D nodes and weights / sqrt(pi); degree 0 = not in use
constexpr int MAX_GH = 8;
struct GhRule {
  int32_t degree, pad;
  double x[MAX_GH], w[MAX_GH];
};
constexpr int MAX_DX = 6, MAX_DU = 2, MAX_N = 8, MAX_DZ = 9, MAX_DZT = 8, MAX_DY = 8;

// Device layout ("AoSoA"): a per-cell record with E fp64 elements per problem is stored as
//   buf[cell][tile][e][lane],  lane = problem % 32, tile = problem / 32
// so that (a) the 32 threads of a warp read element e of their 32 problems as one 256-byte
// coalesced segment and (b) the record of one (cell, tile) is a single contiguous E*256-byte
// block (bulk-copy friendly).
struct KParams {
  // ---- per-cell records
  double* prior;  // [T][ntiles][E_POST][32]  read by the forward sweep
  double* post;   // [T][ntiles][E_POST][32]  written by the backward sweep
  double* latest; // == prior or post: the record the most recent backward sweep wrote (propagate / getters)
  double* filt;   // [T][ntiles][E_FILT][32]  forward -> backward
  double* auxf;   // [T][ntiles][E_AUXF][32]  optional
  double* auxb;   // [T][ntiles][E_AUXB][32]  optional
  double* pf;     // [T][ntiles][E_PF][32]    optional (propagate messages)
  double* term;   // [ntiles][E_TERM][32]     terminal cost-feature moments of the last cell
  double* ric;    // [T][ntiles][E_RIC][32]   Riccati messages (optional): lambda_x3_b, nu_x3_b, lambda_x0_b, nu_x0_b
  // ---- per-problem
  double* x0;      // [ntiles][DX][32]
  double* sig_x0;  // [ntiles][TRI(DX)][32]
  double* alpha;   // [Bpad]
  double* alpha_cell;   // [T][Bpad]   (cells flagged OWN_ALPHA)
  const double* z_cell; // [T][DZ] or [T][ntiles][DZ][32]
  const double* z_term_pp;  // [ntiles][DZT][32] per-problem terminal targets (z_per_problem), else NULL
  const double* envpar; // [ntiles][NP][32]
  int32_t* cell_flags;  // [T]
  const int32_t* cell_index;  // [T]
  double* metrics;      // [I2C_M_COUNT][max_iters][Bpad]
  int32_t* status;      // [Bpad]
  int32_t* info;        // [Bpad]
  // ---- scalars
  int32_t B, Bpad, ntiles, T;
  int32_t n_iter, phases, tau, max_iters;
  int32_t cell_head;  // ring offset: logical cell t lives in slot (t + cell_head) % T (O(1) MPC horizon shift)
  int32_t z_per_problem, qr_diag, has_qf, cov_ctrl;
  int32_t fast_obs;   // cubature rule has zero centre weight and unit weight sum: structured cost-feature moments
  int32_t no_team;    // debugging / A-B: force the one-warp-per-tile kernel
  int32_t stage_meta; // stage the cell targets / flags with the records (latency regime only; set by the launcher)
  int32_t linearize;  // Linearize inference (linear envs only): exact moments instead of sigma points
  double alpha_tol, temp0, dtemp;
  double sf_n, w0_n, wi_n;  // cubature rule in dim n = dx+du (exp_types.py:36-49)
  double sf_x, w0_x, wi_x;  // cubature rule in dim dx
  // ---- shared constants (row-major full matrices; packed lower where noted)
  double QR[MAX_DZ * MAX_DZ], QRinv[MAX_DZ * MAX_DZ];
  double Qf[MAX_DZT * MAX_DZT], Qfinv[MAX_DZT * MAX_DZT];
  double sig_eta[MAX_DX * (MAX_DX + 1) / 2];     // packed lower
  double z_graph[MAX_DZ], z_term[MAX_DZT];
  double sxt[MAX_DX * (MAX_DX + 1) / 2];         // sig_x_terminal, packed lower
  double sxt_inv_mu[MAX_DX];                     // sig_x_terminal^{-1} mu_x_terminal
  double mu_xt[MAX_DX];
  double sxt_logdet;                             // log det sig_x_terminal (KL term)
  GhRule gh;                                     // Gauss-Hermite inference (I2C_INF_GAUSS_HERMITE), else degree 0
  // (new fields go at the END: the offsets of the hot fields above are part of the tuned code generation)
  int32_t group_mode;      // -1 auto, 0 never, 1 always: G-lanes-per-problem kernel (i2c_group.cuh)
  int32_t group_max_tiles; // auto: use it up to this many tiles (0 = built-in policy)
  int32_t minb;            // A/B: resident 128-thread blocks per SM of the small-env throughput variant (0 = default 4)
  int32_t hot;             // the common configuration (Worker<..., HOT>): fast_obs, shared targets, no aux records; 2 = cells with their own alpha allowed
  int32_t* tickets;        // [1 + ntiles]: ticket counter and per-tile finished-iteration counters of em_ticket_kernel
};

// element counts of the records for given dims
struct RecDims {
  int dx, du, n, dz, dzt;
  __host__ __device__ static constexpr int tri(int k) { return k * (k + 1) / 2; }
  __host__ __device__ constexpr int e_post() const { return n + tri(n) + du * dx + du + tri(du); }
  __host__ __device__ constexpr int e_filt() const { return n + tri(n) + dx + tri(dx) + n * dx; }
  __host__ __device__ constexpr int e_auxf() const { return n + tri(n) + dz + tri(dz); }
  __host__ __device__ constexpr int e_auxb() const { return dz + tri(dz) + dx + tri(dx); }
  __host__ __device__ constexpr int e_pf() const { return n + tri(n) + dz + tri(dz) + dx + tri(dx); }
  __host__ __device__ constexpr int e_term() const { return dzt + tri(dzt); }
  __host__ __device__ constexpr int e_ric() const { return 2 * (dx * dx + dx); }
};

// launchers implemented in i2c_kernels.cu
int launch_em(int env, const KParams& p, void* stream);

struct QuadArgs {
  const double* m;   // [ntiles][D][32]
  const double* S;   // [ntiles][TRI(D)][32]
  const double* envpar;
  double* my;        // [ntiles][DY][32]
  double* Sy;        // [ntiles][TRI(DY)][32]
  double* Sxy;       // [ntiles][D*DY][32]
  int32_t* status;   // [Bpad]
  int32_t B, ntiles;
  double sf, w0, wi;
  GhRule gh;         // degree > 0: Gauss-Hermite grid instead of the cubature points
};
int launch_quadrature(int env, int fn, const QuadArgs& a, void* stream);

struct CkfArgs {
  double* x0;        // [ntiles][DX][32]   belief mean, updated in place
  double* sig_x0;    // [ntiles][TRI DX][32]
  const double* y;   // [ntiles][DY][32], or canonical [B][DY] (canonical = 1)
  const double* u;   // [ntiles][DU][32], or canonical [B][DU]
  const double* envpar;
  int32_t* status;
  int32_t B, ntiles;
  int32_t canonical, pad;  // inputs in the caller's [B][d] layout (no pack kernels)
  double sf, w0, wi;
  double sig_eta[MAX_DX * (MAX_DX + 1) / 2];
  double sig_zeta[MAX_DY * (MAX_DY + 1) / 2];
};
int launch_ckf(int env, const CkfArgs& a, void* stream);

// batched closed-loop policy evaluation (canonical layouts, one thread per (problem, roll-out))
struct RolloutArgs {
  const double* x_init;  // [B][R][dx]
  const double* K;       // [B][T][du][dx]
  const double* k;       // [B][T][du]
  const double* sigK;    // [B][T][du][du] or NULL (deterministic policy)
  const double* ex_mu;   // [B][T][dx]     expert gating centre or NULL
  const double* ex_lam;  // [B][T][dx][dx] expert gating precision or NULL
  const double* eta;     // [B][R][T][dx]  process disturbances (already coloured) or NULL -> device RNG
  const double* eps_u;   // [B][R][T][du]  standard normals for action sampling or NULL -> device RNG (if sigK)
  const double* envpar;  // [B][NP] or NULL
  double* xu;            // [B][R][T][n]
  double* z;             // [B][R][T][dz]
  double* z_term;        // [B][R][dzt]
  double* x_final;       // [B][R][dx] state after the last step
  int32_t B, R, T, soft_expert, noise_free;
  uint64_t seed;
  double hard_threshold;
  double chol_eta[MAX_DX * (MAX_DX + 1) / 2];  // lower Cholesky factor of sig_eta (device RNG path)
};
int launch_rollout(int env, const RolloutArgs& a, void* stream);

// chunked parallel-in-time sweep (i2c_scan.cuh); linear environments + Linearize inference only
struct ScanArgs {
  double* fagg;  // [n_chunks][ntiles][E_FAGG][32]  forward chunk aggregates (A, b, C, eta, J)
  double* cin;   // [n_chunks][ntiles][E_MSG][32]   state message entering each chunk from the left
  double* bagg;  // [n_chunks][ntiles][E_BAGG][32]  backward chunk aggregates (E, g, L)
  double* bin;   // [n_chunks][ntiles][E_MSG][32]   smoothed message entering each chunk from the right
  double* part;  // [n_chunks][SCAN_PARTS][Bpad]    per-chunk partial statistics
  double* tail;  // [Bpad]                          terminal trace term of the alpha update
  int32_t n_chunks, chunk, it;
  double temp;
};
int launch_scan(int env, int stage, const KParams& p, const ScanArgs& a, void* stream);
enum { SCAN_FWD_LOCAL = 0, SCAN_FWD_PREFIX, SCAN_FWD_CELLS, SCAN_BWD_LOCAL, SCAN_BWD_SUFFIX, SCAN_BWD_CELLS, SCAN_MSTEP };
enum { SP_ENTX_M = 0, SP_ENTX_E, SP_COST, SP_COST_VAR, SP_TR, SP_ENTU_M, SP_ENTU_E, SCAN_PARTS };


}  // namespace i2c
