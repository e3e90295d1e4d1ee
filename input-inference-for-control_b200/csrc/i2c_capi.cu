// C-ABI of the B200 i2c library (declared in include/i2c_b200.h): handle management, conversion between
// the canonical host layout ([B][T][rows][cols], full symmetric matrices) and the tiled device layout,
// and the launch of the persistent EM kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <utility>
#include <vector>

#include "i2c_types.h"
#include "linalg.cuh"  // fastmath probe (LogAcc, fast_* primitives)

using namespace i2c;

// per-environment launchers, one translation unit each (i2c_env_inst.cu)
namespace i2c {
#define I2C_DECL_ENV(k)                                                   \
  int launch_em_env##k(const KParams& p, void* stream);                   \
  int launch_quad_env##k(int fn, const QuadArgs& a, void* stream);        \
  int launch_ckf_env##k(const CkfArgs& a, void* stream);              \
  int launch_rollout_env##k(const RolloutArgs& a, void* stream);       \
  int launch_scan_env##k(int stage, const KParams& p, const ScanArgs& a, void* stream);
I2C_DECL_ENV(0) I2C_DECL_ENV(1) I2C_DECL_ENV(2) I2C_DECL_ENV(3) I2C_DECL_ENV(4) I2C_DECL_ENV(5) I2C_DECL_ENV(6)
#undef I2C_DECL_ENV

#define I2C_SWITCH_ENV(env, call)   \
  switch (env) {                    \
    case 0: return call(0);         \
    case 1: return call(1);         \
    case 2: return call(2);         \
    case 3: return call(3);         \
    case 4: return call(4);         \
    case 5: return call(5);         \
    case 6: return call(6);         \
  }                                 \
  return -1;

int launch_em(int env, const KParams& p, void* stream) {
#define CALL(k) launch_em_env##k(p, stream)
  I2C_SWITCH_ENV(env, CALL)
#undef CALL
}
int launch_quadrature(int env, int fn, const QuadArgs& a, void* stream) {
#define CALL(k) launch_quad_env##k(fn, a, stream)
  I2C_SWITCH_ENV(env, CALL)
#undef CALL
}
int launch_ckf(int env, const CkfArgs& a, void* stream) {
#define CALL(k) launch_ckf_env##k(a, stream)
  I2C_SWITCH_ENV(env, CALL)
#undef CALL
}
int launch_rollout(int env, const RolloutArgs& a, void* stream) {
#define CALL(k) launch_rollout_env##k(a, stream)
  I2C_SWITCH_ENV(env, CALL)
#undef CALL
}
int launch_scan(int env, int stage, const KParams& p, const ScanArgs& a, void* stream) {
#define CALL(k) launch_scan_env##k(stage, p, a, stream)
  I2C_SWITCH_ENV(env, CALL)
#undef CALL
}
}  // namespace i2c

static thread_local std::string g_err;
static int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(x)                                                                                       \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) return set_err(-100 - (int)e_, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define REQUIRE(c, msg) \
  do {                  \
    if (!(c)) return set_err(-1, msg); \
  } while (0)

// Every entry point that takes a handle runs on the handle's device, whatever device is current in the calling thread
// (a second handle on another GPU, or torch switching devices), and restores the caller's device on return.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct EnvDims {
  int dx, du, dz, dzt, np, dy;
};
static const EnvDims kEnv[I2C_ENV_COUNT] = {
    {2, 1, 3, 2, 8, 0}, {2, 1, 1, 2, 8, 0}, {2, 1, 4, 3, 0, 0}, {2, 1, 1, 1, 0, 0},
    {4, 1, 6, 5, 0, 0}, {6, 1, 9, 8, 0, 0}, {6, 2, 8, 6, 0, 8},
};
static const bool kHasTerm[I2C_ENV_COUNT] = {true, true, true, false, true, true, true};

static inline int tri(int n) { return n * (n + 1) / 2; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct i2c_handle_s {
  i2c_config_t cfg;
  EnvDims d;
  RecDims r;
  int B, Bpad, ntiles, T;
  cudaStream_t stream;
  bool own_ws;
  char* ws;
  size_t ws_bytes;
  // device buffers (inside ws)
  double *recA, *recB, *filt, *auxf, *auxb, *pf, *ric, *term, *x0, *sig_x0, *alpha, *alpha_cell, *z_cell, *z_term_pp, *envpar, *metrics,
      *scratch, *policy_out;
  int32_t *cell_flags_dev, *cell_index_dev, *status, *info, *tickets;
  size_t scratch_elems;
  // host state
  int prior_is_A;   // forward reads recA (1) or recB (0)
  int latest_is_A;  // most recent posterior record
  std::vector<int32_t> flags, index;
  int cell_head;
  int tau;
  double temp, dtemp;
  KParams kp;  // constants + pointers template
  int last_n_iter;
  long long launches;
  cudaEvent_t ev0, ev1, ev_unpacked, ev_copied, ev_mstaged[2], ev_mcopied[2];
  bool mpending[2];
  double* mstage;  // [2][I2C_M_COUNT][Bpad]: staging slots of i2c_get_last_metrics_async
  cudaStream_t copy_stream;
  // overlapped belief upload (i2c_set_initial_state_async): alternate x0 / sig_x0 buffers filled on their own stream while the
  // running sweep reads the current ones; the pointers swap when the upload has been queued
  cudaStream_t up_stream;
  cudaEvent_t ev_uploaded, ev_alt_free;
  double *x0_alt, *sig_x0_alt, *ustage;
  bool belief_swapped;
  bool copy_pending;
  bool problem_set;
  bool has_z_term_pp;
  std::vector<double> mu_u_init_last, sig_u_host;
};

// ----------------------------------------------------------------------------------------- kernels
// canonical [B][nt][rows][cols] (host order) <-> tiled record elements.  kind 0: dense rows x cols at
// `off`; kind 1: symmetric d x d stored packed-lower at `off`; kind 2: sub-block of a packed-lower
// matrix of order `ld` starting at (r0, c0) (used for the u-marginal views).
struct FieldMap {
  int E, off, kind, rows, cols, r0, c0, per_cell;
};

__device__ __forceinline__ int rec_elem(const FieldMap& f, int r, int c) {
  if (f.kind == 0) return f.off + r * f.cols + c;
  int i = r + f.r0, j = c + f.c0;
  return f.off + (i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i);
}

__global__ void unpack_kernel(const double* __restrict__ rec, FieldMap f, int t0, int nt, int T, int head, int B, int ntiles,
                              double* __restrict__ out) {
  const size_t per = (size_t)f.rows * f.cols;
  const size_t total = (size_t)B * nt * per;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % f.cols);
    int r = (int)((i / f.cols) % f.rows);
    int t = (int)((i / per) % nt);
    int b = (int)(i / (per * nt));
    int slot = f.per_cell ? (t0 + t + head) % T : 0;
    size_t src = (((size_t)slot * ntiles + b / TILE) * f.E + rec_elem(f, r, c)) * TILE + (b % TILE);
    out[i] = rec[src];
  }
}

// pack with padding: problems b >= B replicate problem B-1 so that padded lanes do well-posed work
__global__ void pack_kernel(double* __restrict__ rec, FieldMap f, int t0, int nt, int T, int head, int B, int Bpad,
                            int ntiles, const double* __restrict__ in, int bcast_b) {
  const size_t per = (size_t)f.rows * f.cols;
  const size_t total = (size_t)Bpad * nt * per;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % f.cols);
    int r = (int)((i / f.cols) % f.rows);
    int t = (int)((i / per) % nt);
    int b = (int)(i / (per * nt));
    if (f.kind != 0 && (r + f.r0) < (c + f.c0)) continue;  // packed lower: take the lower triangle
    int bs = bcast_b ? 0 : (b < B ? b : B - 1);
    int slot = f.per_cell ? (t0 + t + head) % T : 0;
    size_t dst = (((size_t)slot * ntiles + b / TILE) * f.E + rec_elem(f, r, c)) * TILE + (b % TILE);
    rec[dst] = in[((size_t)bs * nt + t) * per + (size_t)r * f.cols + c];
  }
}

// constructor state of the cells [t0, t0+nt): prior == posterior == N([x0; mu_u], blkdiag(sig_x0, sig_u)), K = 0,
// k = mu_u, sigK = sig_u  (I2cCell.__init__, i2c.py:92-100,135-136)
struct InitArgs {
  int dx, du, n, E, T, head, Bpad, ntiles, bcast_u;
  double sig_u[3];
};
__global__ void init_cells_kernel(double* recA, double* recB, InitArgs a, int t0, int nt, const double* __restrict__ x0,
                                  const double* __restrict__ sig_x0, const double* __restrict__ mu_u /*[B|1][nt][du]*/, int B) {
  const size_t total = (size_t)a.Bpad * nt;
  const int n = a.n, dx = a.dx, du = a.du;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i % a.Bpad), t = (int)(i / a.Bpad);
    int bs = a.bcast_u ? 0 : (b < B ? b : B - 1);
    int slot = (t0 + t + a.head) % a.T;
    size_t base = (((size_t)slot * a.ntiles + b / TILE) * a.E) * TILE + (b % TILE);
    size_t xb = ((size_t)(b / TILE) * dx) * TILE + (b % TILE);
    size_t sb = ((size_t)(b / TILE) * (dx * (dx + 1) / 2)) * TILE + (b % TILE);
    int e = 0;
    for (int k = 0; k < dx; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = x0[xb + (size_t)k * TILE];
    for (int k = 0; k < du; ++k, ++e)
      recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = mu_u[((size_t)bs * nt + t) * du + k];
    for (int r = 0; r < n; ++r)
      for (int c = 0; c <= r; ++c, ++e) {
        double v = 0.0;
        if (r < dx) v = sig_x0[sb + (size_t)(r * (r + 1) / 2 + c) * TILE];
        else if (c >= dx) v = a.sig_u[(r - dx) * (r - dx + 1) / 2 + (c - dx)];
        recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = v;
      }
    for (int k = 0; k < du * dx; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = 0.0;
    for (int k = 0; k < du; ++k, ++e)
      recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = mu_u[((size_t)bs * nt + t) * du + k];
    for (int k = 0; k < du * (du + 1) / 2; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = a.sig_u[k];
  }
}

// fp64 FMA-pipe peak probe (roofline denominator: MEASURED_PEAKS.json has no fp64 figure): 8 independent
// DFMA chains per thread, enough warps to saturate every sub-partition.
__global__ void dfma_peak_kernel(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3,
         a6 = a0 + 6e-3, a7 = a0 + 7e-3;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// rows of the metrics table -> one contiguous [n][B] block (i2c_get_last_metrics_async)
struct MetricIds {
  int32_t v[16];
};
__global__ void gather_metrics_kernel(const double* __restrict__ metrics, MetricIds ids, int n, int it, int max_iters, int B, int Bpad,
                                      double* __restrict__ out) {
  const size_t total = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / B), b = (int)(i % B);
    out[i] = metrics[((size_t)ids.v[m] * max_iters + it) * Bpad + b];
  }
}

// tiny by-value setters: let the MPC step update one cell's metadata without host staging / synchronisation
struct SmallVals {
  double v[16];
};
__global__ void set_vals_kernel(double* dst, SmallVals s, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = s.v[threadIdx.x];
}
__global__ void set_cell_meta_kernel(int32_t* flags, int32_t* index, int slot, int32_t f, int32_t ix) {
  flags[slot] = f;
  index[slot] = ix;
}
__global__ void fill_kernel(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Tail of the MPC step in one launch (policy/mpc.py:166, 174-181): the first planned action = mu_u of the posterior of
// cells[0] is copied to `u_out` ([B][du], caller's layout), then the ring slot of that very cell -- cells.pop(0) -- becomes
// the fresh last cell (constructor state with the initial action mean, its index / target / alpha reset).  The cell flags
// of the whole horizon arrive by value (state after this step's _update_priors): no host staging, no extra copy.
struct SmallInts {
  int32_t v[64];
};
__global__ void mpc_tail_kernel(double* recA, double* recB, const double* latest /* aliases recA or recB */, InitArgs a,
                                const double* __restrict__ x0, const double* __restrict__ sig_x0, SmallVals mu, int B, int slot,
                                int32_t* flags, int32_t* index, SmallInts f, int nflags, double* alpha_cell, double alpha_init,
                                double* z_cell, SmallVals z, int dz, double* __restrict__ u_out) {
  const int n = a.n, dx = a.dx, du = a.du;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.Bpad; b += gridDim.x * blockDim.x) {
    size_t base = (((size_t)slot * a.ntiles + b / TILE) * a.E) * TILE + (b % TILE);
    if (b < B)
      for (int k = 0; k < du; ++k) u_out[(size_t)b * du + k] = latest[base + (size_t)(dx + k) * TILE];
    size_t xb = ((size_t)(b / TILE) * dx) * TILE + (b % TILE);
    size_t sb = ((size_t)(b / TILE) * (dx * (dx + 1) / 2)) * TILE + (b % TILE);
    int e = 0;
    for (int k = 0; k < dx; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = x0[xb + (size_t)k * TILE];
    for (int k = 0; k < du; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = mu.v[k];
    for (int r = 0; r < n; ++r)
      for (int c = 0; c <= r; ++c, ++e) {
        double v = 0.0;
        if (r < dx) v = sig_x0[sb + (size_t)(r * (r + 1) / 2 + c) * TILE];
        else if (c >= dx) v = a.sig_u[(r - dx) * (r - dx + 1) / 2 + (c - dx)];
        recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = v;
      }
    for (int k = 0; k < du * dx; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = 0.0;
    for (int k = 0; k < du; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = mu.v[k];
    for (int k = 0; k < du * (du + 1) / 2; ++k, ++e) recA[base + (size_t)e * TILE] = recB[base + (size_t)e * TILE] = a.sig_u[k];
    alpha_cell[(size_t)slot * a.Bpad + b] = alpha_init;
    if (b == 0) {
      for (int k = 0; k < nflags; ++k) flags[k] = f.v[k];
      index[slot] = 0;
      for (int k = 0; k < dz; ++k) z_cell[(size_t)slot * dz + k] = z.v[k];
    }
  }
}

static inline int nblocks(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// ----------------------------------------------------------------------------------------- layout
struct WsLayout {
  size_t recA, recB, filt, auxf, auxb, pf, ric, term, x0, sig_x0, alpha, alpha_cell, z_cell, z_term_pp, envpar, metrics, scratch, policy_out, flags,
      index, status, info, tickets, mstage, x0_alt, sig_x0_alt, ustage, total, scratch_elems;
};

static WsLayout plan(const i2c_config_t& c, const EnvDims& d) {
  WsLayout w;
  RecDims r{d.dx, d.du, d.dx + d.du, d.dz, d.dzt};
  size_t Bpad = align_up((size_t)c.n_problems, TILE), nt = Bpad / TILE, T = c.horizon;
  size_t off = 0;
  auto take = [&](size_t elems, size_t esz) {
    size_t o = off;
    off = align_up(off + elems * esz, 256);
    return o;
  };
  w.recA = take(T * nt * r.e_post() * TILE, 8);
  w.recB = take(T * nt * r.e_post() * TILE, 8);
  w.filt = take(T * nt * r.e_filt() * TILE, 8);
  w.auxf = take(c.enable_aux ? T * nt * r.e_auxf() * TILE : 0, 8);
  w.auxb = take(c.enable_aux ? T * nt * r.e_auxb() * TILE : 0, 8);
  w.pf = take(c.enable_aux ? T * nt * r.e_pf() * TILE : 0, 8);
  const bool ric = c.enable_aux && c.inference == I2C_INF_LINEARIZE && (c.env == I2C_ENV_LINEAR || c.env == I2C_ENV_LINEAR_MIN_ENERGY);
  w.ric = take(ric ? T * nt * r.e_ric() * TILE : 0, 8);
  w.term = take(nt * r.e_term() * TILE, 8);
  w.x0 = take(nt * d.dx * TILE, 8);
  w.sig_x0 = take(nt * tri(d.dx) * TILE, 8);
  w.alpha = take(Bpad, 8);
  w.alpha_cell = take(T * Bpad, 8);
  w.z_cell = take(c.z_per_problem ? T * nt * d.dz * TILE : T * d.dz, 8);
  w.z_term_pp = take(c.z_per_problem ? nt * d.dzt * TILE : 0, 8);
  w.envpar = take(nt * (d.np > 0 ? d.np : 1) * TILE, 8);
  w.metrics = take((size_t)I2C_M_COUNT * c.max_iters * Bpad, 8);
  size_t n = d.dx + d.du;
  size_t big = n * n;
  if ((size_t)(d.dz * d.dz) > big) big = d.dz * d.dz;
  w.scratch_elems = Bpad * T * big;
  w.scratch = take(w.scratch_elems, 8);
  w.policy_out = take(Bpad * T * (size_t)(d.du * d.dx + d.du + d.du * d.du), 8);
  w.flags = take(T, 4);
  w.index = take(T, 4);
  w.status = take(Bpad, 4);
  w.info = take(Bpad, 4);
  w.mstage = take(2 * (size_t)I2C_M_COUNT * Bpad, 8);
  w.x0_alt = take(nt * d.dx * TILE, 8);
  w.sig_x0_alt = take(nt * tri(d.dx) * TILE, 8);
  w.ustage = take(Bpad * (size_t)(d.dx + d.dx * d.dx), 8);
  w.tickets = take(nt + 1, 4);  // em_ticket_kernel: ticket counter + finished iterations per tile
  w.total = off;
  return w;
}

static int check_cfg(const i2c_config_t* cfg) {
  REQUIRE(cfg != nullptr, "cfg is NULL");
  REQUIRE(cfg->abi_version == I2C_ABI_VERSION, "ABI version mismatch");
  REQUIRE(cfg->env >= 0 && cfg->env < I2C_ENV_COUNT, "unknown env id (no CPU fallback for unregistered envs)");
  REQUIRE(cfg->inference == I2C_INF_CUBATURE || cfg->inference == I2C_INF_LINEARIZE || cfg->inference == I2C_INF_GAUSS_HERMITE,
          "unknown inference kind");
  REQUIRE(cfg->inference != I2C_INF_GAUSS_HERMITE ||
              (cfg->quad_alpha == floor(cfg->quad_alpha) && cfg->quad_alpha >= 1.0 && cfg->quad_alpha <= MAX_GH),
          "Gauss-Hermite degree (passed in quad_alpha) must be an integer in [1, 8]");

  REQUIRE(cfg->n_problems >= 1 && cfg->horizon >= 1 && cfg->horizon < 65536, "bad B or H");
  REQUIRE(cfg->max_iters >= 1, "max_iters must be >= 1");
  REQUIRE(cfg->quad_alpha > 0.0, "cubature alpha must be > 0");
  return 0;
}

extern "C" {

const char* i2c_last_error(void) { return g_err.c_str(); }
const char* i2c_build_info(void) { return "i2c_b200 abi=1 arch=sm_100a fp64 tile=32 " __DATE__ " " __TIME__; }

int i2c_env_dims(int32_t env, int32_t* dx, int32_t* du, int32_t* dz, int32_t* dzt, int32_t* n_par, int32_t* dy) {
  REQUIRE(env >= 0 && env < I2C_ENV_COUNT, "unknown env id");
  const EnvDims& d = kEnv[env];
  if (dx) *dx = d.dx;
  if (du) *du = d.du;
  if (dz) *dz = d.dz;
  if (dzt) *dzt = d.dzt;
  if (n_par) *n_par = d.np;
  if (dy) *dy = d.dy;
  return 0;
}

int i2c_workspace_bytes(const i2c_config_t* cfg, size_t* bytes) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  REQUIRE(bytes != nullptr, "bytes is NULL");
  *bytes = plan(*cfg, kEnv[cfg->env]).total;
  return 0;
}

int i2c_create(const i2c_config_t* cfg, void* workspace_dev, size_t workspace_bytes, void* stream, i2c_handle_t* out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  REQUIRE(out != nullptr, "out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_err(-2, "no CUDA device available: the i2c hot path has no CPU fallback");
  REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
  CUDA_OK(cudaSetDevice(cfg->device));
  i2c_handle_s* h = new i2c_handle_s();
  h->cfg = *cfg;
  h->d = kEnv[cfg->env];
  h->r = RecDims{h->d.dx, h->d.du, h->d.dx + h->d.du, h->d.dz, h->d.dzt};
  h->B = cfg->n_problems;
  h->Bpad = (int)align_up((size_t)h->B, TILE);
  h->ntiles = h->Bpad / TILE;
  h->T = cfg->horizon;
  h->stream = (cudaStream_t)stream;
  WsLayout w = plan(*cfg, h->d);
  if (workspace_dev) {
    if (workspace_bytes < w.total) {
      delete h;
      return set_err(-1, "workspace too small");
    }
    h->ws = (char*)workspace_dev;
    h->own_ws = false;
  } else {
    cudaError_t ce = cudaMalloc((void**)&h->ws, w.total);
    if (ce != cudaSuccess) {
      delete h;
      return set_err(-100 - (int)ce, std::string("cudaMalloc workspace: ") + cudaGetErrorString(ce));
    }
    h->own_ws = true;
  }
  h->ws_bytes = w.total;
  h->recA = (double*)(h->ws + w.recA);
  h->recB = (double*)(h->ws + w.recB);
  h->filt = (double*)(h->ws + w.filt);
  h->auxf = cfg->enable_aux ? (double*)(h->ws + w.auxf) : nullptr;
  h->auxb = cfg->enable_aux ? (double*)(h->ws + w.auxb) : nullptr;
  h->pf = cfg->enable_aux ? (double*)(h->ws + w.pf) : nullptr;
  h->ric = (cfg->enable_aux && cfg->inference == I2C_INF_LINEARIZE && (cfg->env == I2C_ENV_LINEAR || cfg->env == I2C_ENV_LINEAR_MIN_ENERGY))
               ? (double*)(h->ws + w.ric) : nullptr;
  h->term = (double*)(h->ws + w.term);
  h->x0 = (double*)(h->ws + w.x0);
  h->sig_x0 = (double*)(h->ws + w.sig_x0);
  h->alpha = (double*)(h->ws + w.alpha);
  h->alpha_cell = (double*)(h->ws + w.alpha_cell);
  h->z_cell = (double*)(h->ws + w.z_cell);
  h->z_term_pp = cfg->z_per_problem ? (double*)(h->ws + w.z_term_pp) : nullptr;
  h->envpar = (double*)(h->ws + w.envpar);
  h->metrics = (double*)(h->ws + w.metrics);
  h->scratch = (double*)(h->ws + w.scratch);
  h->scratch_elems = w.scratch_elems;
  h->policy_out = (double*)(h->ws + w.policy_out);
  h->cell_flags_dev = (int32_t*)(h->ws + w.flags);
  h->cell_index_dev = (int32_t*)(h->ws + w.index);
  h->status = (int32_t*)(h->ws + w.status);
  h->info = (int32_t*)(h->ws + w.info);
  h->tickets = (int32_t*)(h->ws + w.tickets);
  h->mstage = (double*)(h->ws + w.mstage);
  h->x0_alt = (double*)(h->ws + w.x0_alt);
  h->sig_x0_alt = (double*)(h->ws + w.sig_x0_alt);
  h->ustage = (double*)(h->ws + w.ustage);
  h->belief_swapped = false;
  h->problem_set = false;
  h->launches = 0;
  h->last_n_iter = 0;
  h->cell_head = 0;
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  cudaEventCreateWithFlags(&h->ev_unpacked, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_mstaged[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_mcopied[i], cudaEventDisableTiming);
    h->mpending[i] = false;
  }
  cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_uploaded, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_alt_free, cudaEventDisableTiming);
  cudaEventRecord(h->ev_alt_free, h->stream);
  h->copy_pending = false;
  CUDA_OK(cudaMemsetAsync(h->ws, 0, w.total, h->stream));
  *out = h;
  return 0;
}

int i2c_destroy(i2c_handle_t h) {
  if (!h) return 0;
  DeviceGuard device_guard_(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->copy_stream);
  cudaStreamDestroy(h->copy_stream);
  cudaStreamSynchronize(h->up_stream);
  cudaStreamDestroy(h->up_stream);
  cudaEventDestroy(h->ev_uploaded);
  cudaEventDestroy(h->ev_alt_free);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev_unpacked);
  cudaEventDestroy(h->ev_copied);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(h->ev_mstaged[i]);
    cudaEventDestroy(h->ev_mcopied[i]);
  }
  if (h->own_ws) cudaFree(h->ws);
  delete h;
  return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------- helpers
static int stage_in(i2c_handle_t h, const double* host, size_t elems, double** dev) {
  REQUIRE(elems <= h->scratch_elems, "internal: staging buffer too small");
  CUDA_OK(cudaMemcpyAsync(h->scratch, host, elems * 8, cudaMemcpyHostToDevice, h->stream));
  *dev = h->scratch;
  return 0;
}

// scratch_off / sync: several packs of one API call stage at different offsets of the scratch buffer and share one
// synchronisation (the host buffers are only borrowed for the duration of the call)
static int pack(i2c_handle_t h, double* rec, const FieldMap& f, int t0, int nt, const double* host, bool bcast_b,
                size_t scratch_off = 0, bool sync = true) {
  size_t per = (size_t)f.rows * f.cols;
  size_t elems = (size_t)(bcast_b ? 1 : h->B) * nt * per;
  REQUIRE(scratch_off + elems <= h->scratch_elems, "internal: staging buffer too small");
  double* dev = h->scratch + scratch_off;
  CUDA_OK(cudaMemcpyAsync(dev, host, elems * 8, cudaMemcpyHostToDevice, h->stream));
  size_t total = (size_t)h->Bpad * nt * per;
  pack_kernel<<<nblocks(total), 256, 0, h->stream>>>(rec, f, t0, nt, h->T, f.per_cell ? h->cell_head : 0, h->B, h->Bpad,
                                                     h->ntiles, dev, bcast_b ? 1 : 0);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  // the staging buffer is reused by the next call
  if (sync) CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int unpack(i2c_handle_t h, const double* rec, const FieldMap& f, int t0, int nt, double* host) {
  size_t per = (size_t)f.rows * f.cols;
  size_t total = (size_t)h->B * nt * per;
  REQUIRE(total <= h->scratch_elems, "internal: staging buffer too small");
  unpack_kernel<<<nblocks(total), 256, 0, h->stream>>>(rec, f, t0, nt, h->T, f.per_cell ? h->cell_head : 0, h->B, h->ntiles,
                                                       h->scratch);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(host, h->scratch, total * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

static double* rec_prior(i2c_handle_t h) { return h->prior_is_A ? h->recA : h->recB; }
static double* rec_post(i2c_handle_t h) { return h->prior_is_A ? h->recB : h->recA; }
static double* rec_latest(i2c_handle_t h) { return h->latest_is_A ? h->recA : h->recB; }

static int field_map(i2c_handle_t h, int field, FieldMap* f, double** base) {
  const int dx = h->d.dx, du = h->d.du, n = dx + du, dz = h->d.dz, dzt = h->d.dzt;
  const RecDims& r = h->r;
  const int P_SIG = n, P_K = n + tri(n), P_KK = P_K + du * dx, P_SIGK = P_KK + du;
  const int F_SIG1 = n, F_MU3 = n + tri(n), F_SIG3 = F_MU3 + dx, F_J = F_SIG3 + tri(dx);
  FieldMap m{0, 0, 0, 1, 1, 0, 0, 1};
  switch (field) {
    case I2C_F_MU_XU0_M: m = {r.e_post(), 0, 0, n, 1, 0, 0, 1}; *base = rec_latest(h); break;
    case I2C_F_SIG_XU0_M: m = {r.e_post(), P_SIG, 1, n, n, 0, 0, 1}; *base = rec_latest(h); break;
    case I2C_F_K: m = {r.e_post(), P_K, 0, du, dx, 0, 0, 1}; *base = rec_latest(h); break;
    case I2C_F_KK: m = {r.e_post(), P_KK, 0, du, 1, 0, 0, 1}; *base = rec_latest(h); break;
    case I2C_F_SIGK: m = {r.e_post(), P_SIGK, 1, du, du, 0, 0, 1}; *base = rec_latest(h); break;
    case I2C_F_PRIOR_MU: m = {r.e_post(), 0, 0, n, 1, 0, 0, 1}; *base = rec_prior(h); break;
    case I2C_F_PRIOR_SIG: m = {r.e_post(), P_SIG, 1, n, n, 0, 0, 1}; *base = rec_prior(h); break;
    case I2C_F_PRIOR_K: m = {r.e_post(), P_K, 0, du, dx, 0, 0, 1}; *base = rec_prior(h); break;
    case I2C_F_MU_XU1_F: m = {r.e_filt(), 0, 0, n, 1, 0, 0, 1}; *base = h->filt; break;
    case I2C_F_SIG_XU1_F: m = {r.e_filt(), F_SIG1, 1, n, n, 0, 0, 1}; *base = h->filt; break;
    case I2C_F_MU_X3_F: m = {r.e_filt(), F_MU3, 0, dx, 1, 0, 0, 1}; *base = h->filt; break;
    case I2C_F_SIG_X3_F: m = {r.e_filt(), F_SIG3, 1, dx, dx, 0, 0, 1}; *base = h->filt; break;
    case I2C_F_J_DYN: m = {r.e_filt(), F_J, 0, n, dx, 0, 0, 1}; *base = h->filt; break;
    case I2C_F_MU_XU0_F: m = {r.e_auxf(), 0, 0, n, 1, 0, 0, 1}; *base = h->auxf; break;
    case I2C_F_SIG_XU0_F: m = {r.e_auxf(), n, 1, n, n, 0, 0, 1}; *base = h->auxf; break;
    case I2C_F_MU_Z0_F: m = {r.e_auxf(), n + tri(n), 0, dz, 1, 0, 0, 1}; *base = h->auxf; break;
    case I2C_F_SIG_Z0_F: m = {r.e_auxf(), n + tri(n) + dz, 1, dz, dz, 0, 0, 1}; *base = h->auxf; break;
    case I2C_F_MU_Z0_M: m = {r.e_auxb(), 0, 0, dz, 1, 0, 0, 1}; *base = h->auxb; break;
    case I2C_F_SIG_Z0_M: m = {r.e_auxb(), dz, 1, dz, dz, 0, 0, 1}; *base = h->auxb; break;
    case I2C_F_MU_X3_M: m = {r.e_auxb(), dz + tri(dz), 0, dx, 1, 0, 0, 1}; *base = h->auxb; break;
    case I2C_F_SIG_X3_M: m = {r.e_auxb(), dz + tri(dz) + dx, 1, dx, dx, 0, 0, 1}; *base = h->auxb; break;
    case I2C_F_MU_XU0_PF: m = {r.e_pf(), 0, 0, n, 1, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_SIG_XU0_PF: m = {r.e_pf(), n, 1, n, n, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_MU_Z0_PF: m = {r.e_pf(), n + tri(n), 0, dz, 1, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_SIG_Z0_PF: m = {r.e_pf(), n + tri(n) + dz, 1, dz, dz, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_MU_X3_PF: m = {r.e_pf(), n + tri(n) + dz + tri(dz), 0, dx, 1, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_SIG_X3_PF: m = {r.e_pf(), n + tri(n) + dz + tri(dz) + dx, 1, dx, dx, 0, 0, 1}; *base = h->pf; break;
    case I2C_F_MU_Z3_M: m = {r.e_term(), 0, 0, dzt, 1, 0, 0, 0}; *base = h->term; break;
    case I2C_F_SIG_Z3_M: m = {r.e_term(), dzt, 1, dzt, dzt, 0, 0, 0}; *base = h->term; break;
    case I2C_F_LAMBDA_X3_B: m = {r.e_ric(), 0, 0, dx, dx, 0, 0, 1}; *base = h->ric; break;
    case I2C_F_NU_X3_B: m = {r.e_ric(), dx * dx, 0, dx, 1, 0, 0, 1}; *base = h->ric; break;
    case I2C_F_LAMBDA_X0_B: m = {r.e_ric(), dx * dx + dx, 0, dx, dx, 0, 0, 1}; *base = h->ric; break;
    case I2C_F_NU_X0_B: m = {r.e_ric(), 2 * dx * dx + dx, 0, dx, 1, 0, 0, 1}; *base = h->ric; break;
    default: return set_err(-1, "unknown field id");
  }
  if (*base == nullptr) return set_err(-1, "field needs a handle created with enable_aux=1");
  *f = m;
  return 0;
}

static void full_to_tri(const double* full, int n, double* packed) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) packed[i * (i + 1) / 2 + j] = full[i * n + j];
}

// dense SPD helpers for the handful of host-side constants (sizes <= 9)
static bool host_chol(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    A[j * n + j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / A[j * n + j];
    }
  }
  return true;
}
static bool host_spd_inverse(const double* M, int n, double* out, double* logdet) {
  std::vector<double> L(M, M + n * n);
  if (!host_chol(L, n)) return false;
  double ld = 0.0;
  for (int i = 0; i < n; ++i) ld += 2.0 * log(L[i * n + i]);
  if (logdet) *logdet = ld;
  for (int c = 0; c < n; ++c) {
    std::vector<double> y(n, 0.0);
    for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
      y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * out[k * n + c];
      out[i * n + c] = s / L[i * n + i];
    }
  }
  return true;
}

static void cubature_rule(double a, double b, double k, int dim, double* sf, double* w0, double* wi) {
  // CubatureQuadrature.weights (exp_types.py:40-49); the mean uses weights_sig as well (quirk A.6.1)
  double lam = a * a * (dim + k) - dim;
  *sf = sqrt(dim + lam);
  *wi = 1.0 / (2.0 * (dim + lam));
  *w0 = 2.0 * lam * (*wi) + (1.0 - a * a + b);
}

// Gauss-Hermite nodes / weights (np.polynomial.hermite.hermgauss in exp_types.py:57): roots of the physicists' Hermite
// polynomial H_n by interlacing bisection on the orthonormal recurrence + Newton polish; w_i = 1 / (n p_{n-1}(x_i)^2).
// The weights are returned divided by sqrt(pi), i.e. normalised to sum to one per dimension (exp_types.py:66).
static double hermite_orthonormal(int n, double x, double* pnm1) {
  double p0 = 0.7511255444649425, pm = 0.0;  // pi^(-1/4)
  for (int j = 0; j < n; ++j) {
    double pn = x * sqrt(2.0 / (j + 1)) * p0 - sqrt((double)j / (j + 1)) * pm;
    pm = p0;
    p0 = pn;
  }
  if (pnm1) *pnm1 = pm;
  return p0;
}
static void gauss_hermite_rule(int n, double* x, double* w) {
  std::vector<double> roots, prev;
  for (int k = 1; k <= n; ++k) {
    const double R = sqrt(2.0 * k + 1.0) + 1.0;
    std::vector<double> br;
    br.push_back(-R);
    for (double r : prev) br.push_back(r);
    br.push_back(R);
    roots.assign(k, 0.0);
    for (int i = 0; i < k; ++i) {
      double lo = br[i], hi = br[i + 1];
      double flo = hermite_orthonormal(k, lo, nullptr);
      for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi), fm = hermite_orthonormal(k, mid, nullptr);
        if ((fm > 0) == (flo > 0)) {
          lo = mid;
          flo = fm;
        } else {
          hi = mid;
        }
      }
      double r = 0.5 * (lo + hi);
      for (int it = 0; it < 3; ++it) {  // Newton: p_k' = sqrt(2k) p_{k-1}
        double pm, pk = hermite_orthonormal(k, r, &pm);
        r -= pk / (sqrt(2.0 * k) * pm);
      }
      roots[i] = r;
    }
    prev = roots;
  }
  for (int i = 0; i < n; ++i) {
    x[i] = 0.5 * (roots[i] - roots[n - 1 - i]);  // enforce exact symmetry of the node set (middle node = 0)
    double pm;
    hermite_orthonormal(n, x[i], &pm);
    w[i] = 1.0 / (n * pm * pm) / 1.7724538509055159;
  }
}
static void fill_gh(GhRule* g, int degree) {
  memset(g, 0, sizeof(*g));
  g->degree = degree;
  if (degree > 0) gauss_hermite_rule(degree, g->x, g->w);
}

static int upload_flags(i2c_handle_t h) {
  CUDA_OK(cudaMemcpyAsync(h->cell_flags_dev, h->flags.data(), h->T * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(h->cell_index_dev, h->index.data(), h->T * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" {

static int set_initial_state_impl(i2c_handle_t h, const double* x0, const double* sig_x0, bool sync) {
  REQUIRE(h && x0 && sig_x0, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap fx{h->d.dx, 0, 0, h->d.dx, 1, 0, 0, 0};
  int rc = pack(h, h->x0, fx, 0, 1, x0, false, 0, false);
  if (rc) return rc;
  FieldMap fs{tri(h->d.dx), 0, 1, h->d.dx, h->d.dx, 0, 0, 0};
  return pack(h, h->sig_x0, fs, 0, 1, sig_x0, false, align_up((size_t)h->B * h->d.dx, 32), sync);
}
int i2c_set_initial_state(i2c_handle_t h, const double* x0, const double* sig_x0) {
  return set_initial_state_impl(h, x0, sig_x0, true);
}
// Asynchronous variant: the (pinned) host buffers must stay valid and unchanged until the next synchronising call on this
// handle (i2c_synchronize, any getter).  The belief is uploaded and packed into the ALTERNATE x0 / sig_x0 buffers on the
// handle's upload stream -- concurrently with a sweep that is still reading the current ones -- and the buffers swap roles:
// everything queued on the handle's stream after this call waits for the upload (one event) and sees the new belief.
int i2c_set_initial_state_async(i2c_handle_t h, const double* x0, const double* sig_x0) {
  REQUIRE(h && x0 && sig_x0, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  const int dx = h->d.dx;
  const size_t nx = (size_t)h->B * dx, ns = (size_t)h->B * dx * dx;
  // the alternate buffers were the current ones until the previous swap: wait for the work that used them
  CUDA_OK(cudaStreamWaitEvent(h->up_stream, h->ev_alt_free, 0));
  CUDA_OK(cudaMemcpyAsync(h->ustage, x0, nx * 8, cudaMemcpyHostToDevice, h->up_stream));
  CUDA_OK(cudaMemcpyAsync(h->ustage + nx, sig_x0, ns * 8, cudaMemcpyHostToDevice, h->up_stream));
  FieldMap fx{dx, 0, 0, dx, 1, 0, 0, 0};
  pack_kernel<<<nblocks((size_t)h->Bpad * dx), 256, 0, h->up_stream>>>(h->x0_alt, fx, 0, 1, h->T, 0, h->B, h->Bpad, h->ntiles,
                                                                      h->ustage, 0);
  FieldMap fs{tri(dx), 0, 1, dx, dx, 0, 0, 0};
  pack_kernel<<<nblocks((size_t)h->Bpad * dx * dx), 256, 0, h->up_stream>>>(h->sig_x0_alt, fs, 0, 1, h->T, 0, h->B, h->Bpad,
                                                                           h->ntiles, h->ustage + nx, 0);
  h->launches += 2;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(h->ev_uploaded, h->up_stream));
  // swap: what is queued on the handle's stream so far used the old buffers (they become the alternates, free once that work
  // is done); what follows waits for the upload
  CUDA_OK(cudaEventRecord(h->ev_alt_free, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_uploaded, 0));
  std::swap(h->x0, h->x0_alt);
  std::swap(h->sig_x0, h->sig_x0_alt);
  h->belief_swapped = !h->belief_swapped;
  return 0;
}

// snapshots serialise the workspace with the belief in its primary buffers
static int normalise_belief(i2c_handle_t h) {
  if (!h->belief_swapped) return 0;
  const size_t nt = h->ntiles;
  CUDA_OK(cudaMemcpyAsync(h->x0_alt, h->x0, nt * h->d.dx * TILE * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(h->sig_x0_alt, h->sig_x0, nt * tri(h->d.dx) * TILE * 8, cudaMemcpyDeviceToDevice, h->stream));
  std::swap(h->x0, h->x0_alt);
  std::swap(h->sig_x0, h->sig_x0_alt);
  h->belief_swapped = false;
  CUDA_OK(cudaEventRecord(h->ev_alt_free, h->stream));
  return 0;
}

int i2c_set_initial_state_dev(i2c_handle_t h, const double* x0_dev, const double* sig_x0_dev) {
  REQUIRE(h && x0_dev && sig_x0_dev, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap fx{h->d.dx, 0, 0, h->d.dx, 1, 0, 0, 0};
  size_t total = (size_t)h->Bpad * h->d.dx;
  pack_kernel<<<nblocks(total), 256, 0, h->stream>>>(h->x0, fx, 0, 1, h->T, 0, h->B, h->Bpad, h->ntiles, x0_dev, 0);
  FieldMap fs{tri(h->d.dx), 0, 1, h->d.dx, h->d.dx, 0, 0, 0};
  total = (size_t)h->Bpad * h->d.dx * h->d.dx;
  pack_kernel<<<nblocks(total), 256, 0, h->stream>>>(h->sig_x0, fs, 0, 1, h->T, 0, h->B, h->Bpad, h->ntiles, sig_x0_dev, 0);
  h->launches += 2;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int i2c_get_initial_state(i2c_handle_t h, double* x0, double* sig_x0) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  int rc = 0;
  if (x0) {
    FieldMap fx{h->d.dx, 0, 0, h->d.dx, 1, 0, 0, 0};
    rc = unpack(h, h->x0, fx, 0, 1, x0);
    if (rc) return rc;
  }
  if (sig_x0) {
    FieldMap fs{tri(h->d.dx), 0, 1, h->d.dx, h->d.dx, 0, 0, 0};
    rc = unpack(h, h->sig_x0, fs, 0, 1, sig_x0);
  }
  return rc;
}

int i2c_set_problem(i2c_handle_t h, const double* x0, const double* sig_x0, const double* sig_eta, const double* mu_u,
                    const double* sig_u, const double* QR, const double* Qf, const double* z, const double* z_graph,
                    const double* z_term, const double* alpha0, double alpha_update_tol, const double* mu_x_term,
                    const double* sig_x_term, double dtemp, const double* env_par) {
  REQUIRE(h && x0 && sig_x0 && sig_eta && mu_u && sig_u && QR && z && z_graph && alpha0, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  const int dx = h->d.dx, du = h->d.du, n = dx + du, dz = h->d.dz, dzt = h->d.dzt, T = h->T;
  REQUIRE(h->d.np == 0 || env_par != nullptr, "this env needs per-problem parameters (env_par)");
  REQUIRE((mu_x_term == nullptr) == (sig_x_term == nullptr), "covariance control needs both mu_x_term and sig_x_term");
  KParams& kp = h->kp;
  memset(&kp, 0, sizeof(kp));
  // ---- constants
  kp.qr_diag = 1;
  for (int a = 0; a < dz; ++a)
    for (int b = 0; b < dz; ++b) {
      kp.QR[a * dz + b] = QR[a * dz + b];
      if (a != b && QR[a * dz + b] != 0.0) kp.qr_diag = 0;
    }
  REQUIRE(host_spd_inverse(QR, dz, kp.QRinv, nullptr), "QR (block_diag(Q, R)) must be positive definite");
  kp.has_qf = (Qf != nullptr && kHasTerm[h->cfg.env]) ? 1 : 0;
  if (Qf != nullptr) {
    REQUIRE(z_term != nullptr, "Qf given without z_term");
    for (int a = 0; a < dzt * dzt; ++a) kp.Qf[a] = Qf[a];
    REQUIRE(host_spd_inverse(Qf, dzt, kp.Qfinv, nullptr), "Qf must be positive definite");
    for (int a = 0; a < dzt; ++a) kp.z_term[a] = z_term[a];
  }
  h->has_z_term_pp = false;
  full_to_tri(sig_eta, dx, kp.sig_eta);
  for (int a = 0; a < dz; ++a) kp.z_graph[a] = z_graph[a];
  kp.cov_ctrl = sig_x_term != nullptr;
  if (kp.cov_ctrl) {
    full_to_tri(sig_x_term, dx, kp.sxt);
    std::vector<double> inv(dx * dx);
    REQUIRE(host_spd_inverse(sig_x_term, dx, inv.data(), &kp.sxt_logdet), "sig_x_terminal must be positive definite");
    for (int i = 0; i < dx; ++i) {
      kp.mu_xt[i] = mu_x_term[i];
      double s = 0.0;
      for (int k = 0; k < dx; ++k) s += inv[i * dx + k] * mu_x_term[k];
      kp.sxt_inv_mu[i] = s;
    }
  }
  kp.alpha_tol = alpha_update_tol;
  cubature_rule(h->cfg.quad_alpha, h->cfg.quad_beta, h->cfg.quad_kappa, n, &kp.sf_n, &kp.w0_n, &kp.wi_n);
  cubature_rule(h->cfg.quad_alpha, h->cfg.quad_beta, h->cfg.quad_kappa, dx, &kp.sf_x, &kp.w0_x, &kp.wi_x);
  fill_gh(&kp.gh, h->cfg.inference == I2C_INF_GAUSS_HERMITE ? (int)h->cfg.quad_alpha : 0);
  if (kp.gh.degree > 0) cubature_rule(1.0, 0.0, 0.0, n, &kp.sf_n, &kp.w0_n, &kp.wi_n), cubature_rule(1.0, 0.0, 0.0, dx, &kp.sf_x, &kp.w0_x, &kp.wi_x);
  kp.fast_obs = kp.gh.degree == 0 && kp.w0_n == 0.0 && kp.w0_x == 0.0 && fabs(2.0 * n * kp.wi_n - 1.0) < 1e-15 &&
                fabs(2.0 * dx * kp.wi_x - 1.0) < 1e-15 && fabs(kp.sf_n * kp.sf_n - n) < 1e-12 && getenv("I2C_B200_GENERIC_OBS") == nullptr;
  // ---- graph state
  h->cell_head = 0;
  h->flags.assign(T, I2C_CELL_INDEPENDENT | I2C_CELL_EXPERT);
  h->flags[T - 1] |= I2C_CELL_TERMINAL;
  h->index.resize(T);
  for (int t = 0; t < T; ++t) h->index[t] = t;
  h->tau = T - 1;
  h->temp = 1.0;
  h->dtemp = dtemp;
  h->prior_is_A = 1;
  h->latest_is_A = 0;
  int rc = upload_flags(h);
  if (rc) return rc;
  // ---- per-problem data
  rc = i2c_set_initial_state(h, x0, sig_x0);
  if (rc) return rc;
  {
    FieldMap fa{1, 0, 0, 1, 1, 0, 0, 0};
    // alpha is [Bpad] flat == tiles of 32 with E = 1
    rc = pack(h, h->alpha, fa, 0, 1, alpha0, false);
    if (rc) return rc;
  }
  if (h->d.np > 0) {
    FieldMap fp{h->d.np, 0, 0, h->d.np, 1, 0, 0, 0};
    rc = pack(h, h->envpar, fp, 0, 1, env_par, false);
    if (rc) return rc;
  }
  if (h->cfg.z_per_problem) {
    FieldMap fz{dz, 0, 0, dz, 1, 0, 0, 1};
    rc = pack(h, h->z_cell, fz, 0, T, z, false);
    if (rc) return rc;
    if (Qf != nullptr) {  // z_term is [B][dzt] when targets are per problem
      FieldMap ft{dzt, 0, 0, dzt, 1, 0, 0, 0};
      rc = pack(h, h->z_term_pp, ft, 0, 1, z_term, false);
      if (rc) return rc;
      h->has_z_term_pp = true;
    }
  } else {
    CUDA_OK(cudaMemcpyAsync(h->z_cell, z, (size_t)T * dz * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  // ---- cells: constructor state
  {
    double* dev;
    rc = stage_in(h, mu_u, (size_t)h->B * T * du, &dev);
    if (rc) return rc;
    InitArgs ia{dx, du, n, h->r.e_post(), T, 0, h->Bpad, h->ntiles, 0, {0, 0, 0}};
    full_to_tri(sig_u, du, ia.sig_u);
    h->sig_u_host.assign(ia.sig_u, ia.sig_u + tri(du));
    init_cells_kernel<<<nblocks((size_t)h->Bpad * T), 256, 0, h->stream>>>(h->recA, h->recB, ia, 0, T, h->x0, h->sig_x0, dev,
                                                                          h->B);
    h->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemsetAsync(h->status, 0, (size_t)h->Bpad * 4, h->stream));
    CUDA_OK(cudaMemsetAsync(h->info, 0, (size_t)h->Bpad * 4, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  h->problem_set = true;
  return 0;
}

int i2c_set_cell_flags(i2c_handle_t h, const int32_t* flags) {
  REQUIRE(h && flags, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  for (int t = 0; t < h->T; ++t) h->flags[(t + h->cell_head) % h->T] = flags[t];
  return upload_flags(h);
}
int i2c_get_cell_flags(i2c_handle_t h, int32_t* flags) {
  REQUIRE(h && flags, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  for (int t = 0; t < h->T; ++t) flags[t] = h->flags[(t + h->cell_head) % h->T];
  return 0;
}
int i2c_set_cell_index(i2c_handle_t h, const int32_t* index) {
  REQUIRE(h && index, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  for (int t = 0; t < h->T; ++t) h->index[(t + h->cell_head) % h->T] = index[t];
  return upload_flags(h);
}
int i2c_set_tau(i2c_handle_t h, int32_t tau) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  h->tau = tau;
  return 0;
}
int i2c_set_cell_targets(i2c_handle_t h, const double* z) {
  REQUIRE(h && z, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  const int dz = h->d.dz, T = h->T;
  if (h->cfg.z_per_problem) {
    FieldMap fz{dz, 0, 0, dz, 1, 0, 0, 1};
    return pack(h, h->z_cell, fz, 0, T, z, false);
  }
  for (int t = 0; t < T; ++t) {
    int slot = (t + h->cell_head) % T;
    CUDA_OK(cudaMemcpyAsync(h->z_cell + (size_t)slot * dz, z + (size_t)t * dz, (size_t)dz * 8, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_set_alpha(i2c_handle_t h, const double* alpha) {
  REQUIRE(h && alpha, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap fa{1, 0, 0, 1, 1, 0, 0, 0};
  int rc = pack(h, h->alpha, fa, 0, 1, alpha, false);
  if (rc) return rc;
  for (auto& f : h->flags) f &= ~I2C_CELL_OWN_ALPHA;  // update_xi pushes sig_xi to all current cells
  return upload_flags(h);
}
int i2c_get_alpha(i2c_handle_t h, double* alpha) {
  REQUIRE(h && alpha, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  CUDA_OK(cudaMemcpyAsync(alpha, h->alpha, (size_t)h->B * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}
int i2c_set_temp(i2c_handle_t h, double temp) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  h->temp = temp;
  return 0;
}
int i2c_get_temp(i2c_handle_t h, double* temp) {
  REQUIRE(h && temp, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  *temp = h->temp;
  return 0;
}

static KParams run_params(i2c_handle_t h, int32_t n_iter, int32_t phases) {
  KParams kp = h->kp;
  kp.prior = rec_prior(h);
  kp.post = rec_post(h);
  kp.latest = rec_latest(h);
  kp.filt = h->filt;
  kp.auxf = h->auxf;
  kp.auxb = h->auxb;
  kp.pf = h->pf;
  kp.ric = h->ric;
  kp.linearize = h->cfg.inference == I2C_INF_LINEARIZE;
  kp.no_team = getenv("I2C_B200_NO_TEAM") != nullptr;
  {
    const char* gm = getenv("I2C_B200_GROUP");
    kp.group_mode = gm ? atoi(gm) : -1;
    const char* mb = getenv("I2C_B200_MINB");
    kp.minb = mb ? atoi(mb) : 0;
    const char* gt = getenv("I2C_B200_GROUP_MAX_TILES");
    kp.group_max_tiles = gt ? atoi(gt) : 0;
  }
  kp.term = h->term;
  kp.x0 = h->x0;
  kp.sig_x0 = h->sig_x0;
  kp.alpha = h->alpha;
  kp.alpha_cell = h->alpha_cell;
  kp.z_cell = h->z_cell;
  kp.z_term_pp = h->has_z_term_pp ? h->z_term_pp : nullptr;
  kp.envpar = h->envpar;
  kp.cell_flags = h->cell_flags_dev;
  kp.cell_index = h->cell_index_dev;
  kp.metrics = h->metrics;
  kp.status = h->status;
  kp.info = h->info;
  kp.tickets = h->tickets;
  kp.B = h->B;
  kp.Bpad = h->Bpad;
  kp.ntiles = h->ntiles;
  kp.T = h->T;
  kp.n_iter = n_iter;
  kp.phases = phases;
  kp.tau = h->tau;
  kp.max_iters = h->cfg.max_iters;
  kp.cell_head = h->cell_head;
  kp.z_per_problem = h->cfg.z_per_problem;
  kp.temp0 = h->temp;
  kp.dtemp = h->dtemp;
  {
    bool own = false;
    for (int f : h->flags) own = own || (f & I2C_CELL_OWN_ALPHA);
    const bool hot = kp.fast_obs && !kp.z_per_problem && !(phases & I2C_PH_STORE_AUX) && !kp.linearize && kp.gh.degree == 0 &&
                     (size_t)h->T * (h->d.dz + 1) * 8 <= 64 * 1024 /* shared-memory table of the horizon */ &&
                     getenv("I2C_B200_NO_HOT") == nullptr;
    // cells with their own alpha (MPC horizon shift): the HOT = 2 instantiation, built for the environments with a measurement
    // model (the partially observed MPC loop); elsewhere the generic kernels
    kp.hot = !hot ? 0 : (!own ? 1 : (h->d.dy > 0 ? 2 : 0));
  }
  return kp;
}

// upload = false: the caller takes care of the device copy of the cell flags (i2c_mpc_step: they ride in its tail kernel)
static int run_core(i2c_handle_t h, int32_t n_iter, int32_t phases, bool upload) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(h->problem_set, "i2c_set_problem has not been called");
  REQUIRE(n_iter >= 1 && n_iter <= h->cfg.max_iters, "n_iter must be in [1, max_iters]");
  REQUIRE(!(phases & I2C_PH_STORE_AUX) || h->cfg.enable_aux, "I2C_PH_STORE_AUX needs enable_aux=1");
  REQUIRE(!(phases & I2C_PH_CALIBRATE) || (phases & I2C_PH_PROPAGATE), "CALIBRATE needs PROPAGATE");
  REQUIRE(!(phases & I2C_PH_RICCATI) || (h->ric != nullptr), "RICCATI needs Linearize inference on a linear environment and enable_aux=1");
  KParams kp = run_params(h, n_iter, phases);
  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  int rc = launch_em(h->cfg.env, kp, (void*)h->stream);
  if (rc != 0) return set_err(-100 - rc, std::string("EM kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  h->launches++;
  h->last_n_iter = n_iter;
  // ---- mirror the state machine the kernel just executed
  bool flip = false, alpha_pushed = false;
  for (int it = 0; it < n_iter; ++it) {
    if (phases & I2C_PH_BACKWARD) {
      h->latest_is_A = h->prior_is_A ? 0 : 1;  // backward wrote `post`
      if (h->kp.cov_ctrl) h->temp += h->dtemp;
    }
    if (phases & I2C_PH_UPDATE_PRIORS) {
      if (h->latest_is_A != h->prior_is_A) h->prior_is_A = h->latest_is_A;  // swap prior/post
      flip = true;
    }
    if (phases & (I2C_PH_MSTEP | I2C_PH_CALIBRATE)) alpha_pushed = true;
  }
  if (flip || alpha_pushed) {
    bool changed = false;
    for (int s = 0; s < h->T; ++s) {
      const int32_t before = h->flags[s];
      if (flip && h->tau > 0 && h->index[s] <= h->tau) h->flags[s] &= ~I2C_CELL_INDEPENDENT;
      if (alpha_pushed) h->flags[s] &= ~I2C_CELL_OWN_ALPHA;
      changed = changed || h->flags[s] != before;
    }
    // (a pageable H2D copy serialises with the stream: only when a flag really changed, i.e. after the first iteration)
    if (changed && upload) CUDA_OK(cudaMemcpyAsync(h->cell_flags_dev, h->flags.data(), h->T * 4, cudaMemcpyHostToDevice, h->stream));
  }
  return 0;
}

int i2c_run(i2c_handle_t h, int32_t n_iter, int32_t phases) { return run_core(h, n_iter, phases, true); }

// Parallel-in-time variant of i2c_run (csrc/i2c_scan.cuh): the horizon is cut into chunks of `chunk_cells` cells that are
// processed concurrently; exact for Linearize inference on the linear environments.  Same records, metrics and state
// machine as i2c_run(FORWARD | BACKWARD [| MSTEP] [| UPDATE_PRIORS] [| STORE_AUX]).
int i2c_run_scan(i2c_handle_t h, int32_t n_iter, int32_t phases, int32_t chunk_cells) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(h->problem_set, "i2c_set_problem has not been called");
  REQUIRE(n_iter >= 1 && n_iter <= h->cfg.max_iters, "n_iter must be in [1, max_iters]");
  REQUIRE(!(phases & I2C_PH_STORE_AUX) || h->cfg.enable_aux, "I2C_PH_STORE_AUX needs enable_aux=1");
  REQUIRE(h->cfg.inference == I2C_INF_LINEARIZE && (h->cfg.env == I2C_ENV_LINEAR || h->cfg.env == I2C_ENV_LINEAR_MIN_ENERGY),
          "the parallel-in-time sweep is exact only for Linearize inference on the linear environments (no fallback)");
  const int32_t allowed = I2C_PH_FORWARD | I2C_PH_BACKWARD | I2C_PH_MSTEP | I2C_PH_UPDATE_PRIORS | I2C_PH_STORE_AUX;
  REQUIRE((phases & ~allowed) == 0, "i2c_run_scan supports FORWARD, BACKWARD, MSTEP, UPDATE_PRIORS, STORE_AUX");
  REQUIRE((phases & I2C_PH_BACKWARD) || !(phases & I2C_PH_MSTEP), "MSTEP needs BACKWARD");
  REQUIRE((phases & I2C_PH_FORWARD) || !(phases & I2C_PH_MSTEP), "MSTEP needs FORWARD");
  const int T = h->T, dx = h->d.dx;
  REQUIRE(chunk_cells >= 8 && chunk_cells <= T, "chunk_cells must be in [8, horizon]");
  const int n_chunks = (T + chunk_cells - 1) / chunk_cells;
  ScanArgs a{};
  {
    size_t per = (size_t)h->ntiles * TILE, off = 0;
    auto take = [&](size_t e) {
      double* q = h->scratch + off;
      off += (size_t)n_chunks * e * per;
      return q;
    };
    a.fagg = take(3 * dx * dx + 2 * dx);
    a.cin = take(dx + tri(dx));
    a.bagg = take(2 * dx * dx + dx);
    a.bin = take(dx + tri(dx));
    a.part = take(SCAN_PARTS);
    a.tail = h->scratch + off;
    off += per;
    REQUIRE(off <= h->scratch_elems, "internal: scan scratch does not fit");
  }
  a.n_chunks = n_chunks;
  a.chunk = chunk_cells;
  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  for (int it = 0; it < n_iter; ++it) {
    for (int s = 0; s < T; ++s)
      REQUIRE((h->flags[s] & I2C_CELL_INDEPENDENT) || !(h->flags[s] & I2C_CELL_EXPERT),
              "parallel-in-time sweep: a feedback cell with use_expert_controller weights K by a pdf ratio of the incoming "
              "message (i2c.py:259-265), which is not linear-Gaussian; clear I2C_CELL_EXPERT or use i2c_run");
    KParams kp = run_params(h, 1, phases);
    a.it = it;
    a.temp = h->temp;
    auto go = [&](int stage) -> int {
      int rc = launch_scan(h->cfg.env, stage, kp, a, (void*)h->stream);
      h->launches++;
      return rc;
    };
    int rc = 0;
    if (phases & I2C_PH_FORWARD) {
      if (n_chunks > 1 && !rc) rc = go(SCAN_FWD_LOCAL);
      if (!rc) rc = go(SCAN_FWD_PREFIX);
      if (!rc) rc = go(SCAN_FWD_CELLS);
    }
    if (phases & I2C_PH_BACKWARD) {
      if (n_chunks > 1 && !rc) rc = go(SCAN_BWD_LOCAL);
      if (!rc) rc = go(SCAN_BWD_SUFFIX);
      if (!rc) rc = go(SCAN_BWD_CELLS);
      h->latest_is_A = h->prior_is_A ? 0 : 1;  // backward wrote `post`
      if (h->kp.cov_ctrl) h->temp += h->dtemp;
    }
    if ((phases & I2C_PH_MSTEP) && !rc) rc = go(SCAN_MSTEP);
    if (rc != 0) return set_err(-100 - rc, std::string("scan kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    bool flags_dirty = false;
    if (phases & I2C_PH_UPDATE_PRIORS) {
      if (h->latest_is_A != h->prior_is_A) h->prior_is_A = h->latest_is_A;  // swap prior/post
      for (int s = 0; s < T; ++s)
        if (h->tau > 0 && h->index[s] <= h->tau && (h->flags[s] & I2C_CELL_INDEPENDENT)) {
          h->flags[s] &= ~I2C_CELL_INDEPENDENT;
          flags_dirty = true;
        }
    }
    if (phases & I2C_PH_MSTEP)
      for (int s = 0; s < T; ++s)
        if (h->flags[s] & I2C_CELL_OWN_ALPHA) {
          h->flags[s] &= ~I2C_CELL_OWN_ALPHA;
          flags_dirty = true;
        }
    if (flags_dirty) CUDA_OK(cudaMemcpyAsync(h->cell_flags_dev, h->flags.data(), T * 4, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  h->last_n_iter = n_iter;
  return 0;
}

int i2c_synchronize(i2c_handle_t h) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_get_metric(i2c_handle_t h, int32_t metric, double* out, int32_t n_iter) {
  REQUIRE(h && out, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(metric >= 0 && metric < I2C_M_COUNT, "unknown metric id");
  REQUIRE(n_iter >= 1 && n_iter <= h->cfg.max_iters, "bad n_iter");
  const double* src = h->metrics + (size_t)metric * h->cfg.max_iters * h->Bpad;
  CUDA_OK(cudaMemcpy2DAsync(out, (size_t)h->B * 8, src, (size_t)h->Bpad * 8, (size_t)h->B * 8, n_iter, cudaMemcpyDeviceToHost,
                            h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_get_status(i2c_handle_t h, int32_t* status, int32_t* info) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  if (status) CUDA_OK(cudaMemcpyAsync(status, h->status, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
  if (info) CUDA_OK(cudaMemcpyAsync(info, h->info, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_clear_status(i2c_handle_t h) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  CUDA_OK(cudaMemsetAsync(h->status, 0, (size_t)h->Bpad * 4, h->stream));
  CUDA_OK(cudaMemsetAsync(h->info, 0, (size_t)h->Bpad * 4, h->stream));
  return 0;
}

int i2c_field_shape(i2c_handle_t h, int32_t field, int32_t* rows, int32_t* cols) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap f;
  double* base = nullptr;
  // shape queries must work without aux buffers
  bool had = h->cfg.enable_aux;
  double* dummy = (double*)1;
  double *a0 = h->auxf, *a1 = h->auxb, *a2 = h->pf, *a3 = h->ric;
  if (!had) h->auxf = h->auxb = h->pf = dummy;
  if (!h->ric) h->ric = dummy;
  int rc = field_map(h, field, &f, &base);
  h->auxf = a0, h->auxb = a1, h->pf = a2, h->ric = a3;
  if (rc) return rc;
  if (rows) *rows = f.rows;
  if (cols) *cols = f.cols;
  return 0;
}

int i2c_get_field(i2c_handle_t h, int32_t field, int32_t t0, int32_t t1, double* out) {
  REQUIRE(h && out, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap f;
  double* base;
  int rc = field_map(h, field, &f, &base);
  if (rc) return rc;
  if (!f.per_cell) {
    t0 = 0;
    t1 = 1;
  }
  REQUIRE(t0 >= 0 && t1 <= h->T && t0 < t1, "bad cell range");
  return unpack(h, base, f, t0, t1 - t0, out);
}

int i2c_set_field(i2c_handle_t h, int32_t field, int32_t t0, int32_t t1, const double* in) {
  REQUIRE(h && in, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  FieldMap f;
  double* base;
  int rc = field_map(h, field, &f, &base);
  if (rc) return rc;
  if (!f.per_cell) {
    t0 = 0;
    t1 = 1;
  }
  REQUIRE(t0 >= 0 && t1 <= h->T && t0 < t1, "bad cell range");
  return pack(h, base, f, t0, t1 - t0, in, false);
}

int i2c_get_policy(i2c_handle_t h, double* K, double* k, double* sigK) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  int rc = 0;
  if (K && (rc = i2c_get_field(h, I2C_F_K, 0, h->T, K))) return rc;
  if (k && (rc = i2c_get_field(h, I2C_F_KK, 0, h->T, k))) return rc;
  if (sigK && (rc = i2c_get_field(h, I2C_F_SIGK, 0, h->T, sigK))) return rc;
  return 0;
}

int i2c_get_policy_dev(i2c_handle_t h, double* K_dev, double* k_dev, double* sigK_dev) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  const int fields[3] = {I2C_F_K, I2C_F_KK, I2C_F_SIGK};
  double* outs[3] = {K_dev, k_dev, sigK_dev};
  for (int i = 0; i < 3; ++i) {
    if (!outs[i]) continue;
    FieldMap f;
    double* base;
    int rc = field_map(h, fields[i], &f, &base);
    if (rc) return rc;
    size_t total = (size_t)h->B * h->T * f.rows * f.cols;
    unpack_kernel<<<nblocks(total), 256, 0, h->stream>>>(base, f, 0, h->T, h->T, h->cell_head, h->B, h->ntiles, outs[i]);
    h->launches++;
    CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int i2c_get_metrics(i2c_handle_t h, const int32_t* metrics, int32_t n_metrics, double* out, int32_t n_iter) {
  REQUIRE(h && metrics && out && n_metrics >= 1, "bad argument");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(n_iter >= 1 && n_iter <= h->cfg.max_iters, "bad n_iter");
  for (int i = 0; i < n_metrics; ++i) {
    REQUIRE(metrics[i] >= 0 && metrics[i] < I2C_M_COUNT, "unknown metric id");
    const double* src = h->metrics + (size_t)metrics[i] * h->cfg.max_iters * h->Bpad;
    CUDA_OK(cudaMemcpy2DAsync(out + (size_t)i * n_iter * h->B, (size_t)h->B * 8, src, (size_t)h->Bpad * 8, (size_t)h->B * 8, n_iter,
                              cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}
// Pipelined read of a step's result: the metrics of the LAST iteration of the most recent i2c_run are gathered into staging
// slot `slot` (0 / 1) on the handle's stream and copied to `out` ([n_metrics][B], page-locked) on the copy stream; the
// caller may queue the next step at once and collects the numbers later with i2c_metrics_wait(slot).  With two slots the
// host stays one step ahead of the device: no per-step drain of the stream (the synchronous i2c_get_metrics costs one).
int i2c_get_last_metrics_async(i2c_handle_t h, const int32_t* metrics, int32_t n_metrics, double* out, int32_t slot) {
  REQUIRE(h && metrics && out && n_metrics >= 1 && n_metrics <= I2C_M_COUNT && n_metrics <= 16, "bad argument");
  REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  REQUIRE(h->last_n_iter >= 1, "no i2c_run yet");
  DeviceGuard device_guard_(h->cfg.device);
  MetricIds ids;
  for (int i = 0; i < n_metrics; ++i) {
    REQUIRE(metrics[i] >= 0 && metrics[i] < I2C_M_COUNT, "unknown metric id");
    ids.v[i] = metrics[i];
  }
  // the previous copy out of this slot must have drained before the slot is overwritten
  if (h->mpending[slot]) CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_mcopied[slot], 0));
  double* st = h->mstage + (size_t)slot * I2C_M_COUNT * h->Bpad;
  const size_t total = (size_t)n_metrics * h->B;
  gather_metrics_kernel<<<nblocks(total), 256, 0, h->stream>>>(h->metrics, ids, n_metrics, h->last_n_iter - 1, h->cfg.max_iters, h->B,
                                                             h->Bpad, st);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(h->ev_mstaged[slot], h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_mstaged[slot], 0));
  CUDA_OK(cudaMemcpyAsync(out, st, total * 8, cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_OK(cudaEventRecord(h->ev_mcopied[slot], h->copy_stream));
  h->mpending[slot] = true;
  return 0;
}

int i2c_metrics_wait(i2c_handle_t h, int32_t slot) {
  REQUIRE(h, "NULL handle");
  REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  DeviceGuard device_guard_(h->cfg.device);
  if (h->mpending[slot]) CUDA_OK(cudaEventSynchronize(h->ev_mcopied[slot]));
  h->mpending[slot] = false;
  return 0;
}

int i2c_get_policy_async(i2c_handle_t h, double* K, double* k, double* sigK) {
  REQUIRE(h && K && k && sigK, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  const size_t BT = (size_t)h->B * h->T, nK = BT * h->d.du * h->d.dx, nk = BT * h->d.du, ns = BT * h->d.du * h->d.du;
  // the previous asynchronous copy must have drained before its device staging buffer is overwritten
  if (h->copy_pending) CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_copied, 0));
  double* outs[3] = {h->policy_out, h->policy_out + nK, h->policy_out + nK + nk};
  const int fields[3] = {I2C_F_K, I2C_F_KK, I2C_F_SIGK};
  for (int i = 0; i < 3; ++i) {
    FieldMap f;
    double* base;
    int rc = field_map(h, fields[i], &f, &base);
    if (rc) return rc;
    size_t total = BT * f.rows * f.cols;
    unpack_kernel<<<nblocks(total), 256, 0, h->stream>>>(base, f, 0, h->T, h->T, h->cell_head, h->B, h->ntiles, outs[i]);
    h->launches++;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(h->ev_unpacked, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_unpacked, 0));
  CUDA_OK(cudaMemcpyAsync(K, outs[0], nK * 8, cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_OK(cudaMemcpyAsync(k, outs[1], nk * 8, cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_OK(cudaMemcpyAsync(sigK, outs[2], ns * 8, cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_OK(cudaEventRecord(h->ev_copied, h->copy_stream));
  h->copy_pending = true;
  return 0;
}

int i2c_copy_wait(i2c_handle_t h) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  if (h->copy_pending) CUDA_OK(cudaEventSynchronize(h->ev_copied));
  h->copy_pending = false;
  return 0;
}

int i2c_get_first_action(i2c_handle_t h, double* mu_u, double* sig_u) {
  REQUIRE(h, "NULL handle");
  DeviceGuard device_guard_(h->cfg.device);
  const int dx = h->d.dx, du = h->d.du, n = dx + du;
  int rc = 0;
  if (mu_u) {
    FieldMap f{h->r.e_post(), dx, 0, du, 1, 0, 0, 1};
    rc = unpack(h, rec_latest(h), f, 0, 1, mu_u);
    if (rc) return rc;
  }
  if (sig_u) {
    FieldMap f{h->r.e_post(), n, 2, du, du, dx, dx, 1};
    rc = unpack(h, rec_latest(h), f, 0, 1, sig_u);
  }
  return rc;
}

int i2c_shift_horizon(i2c_handle_t h, const double* z_new, const double* mu_u_init, double alpha_init) {
  REQUIRE(h && z_new && mu_u_init, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  const int dx = h->d.dx, du = h->d.du, n = dx + du, dz = h->d.dz, T = h->T;
  // cells.pop(0): advance the ring; the freed slot becomes the new last cell
  h->cell_head = (h->cell_head + 1) % T;
  const int slot = (T - 1 + h->cell_head) % T;
  h->flags[slot] = I2C_CELL_INDEPENDENT | I2C_CELL_EXPERT | I2C_CELL_OWN_ALPHA;
  h->index[slot] = 0;
  int rc = upload_flags(h);
  if (rc) return rc;
  // deepcopy(cell_init): constructor-state records (policy/mpc.py:174-176)
  double* dev;
  rc = stage_in(h, mu_u_init, du, &dev);
  if (rc) return rc;
  InitArgs ia{dx, du, n, h->r.e_post(), T, h->cell_head, h->Bpad, h->ntiles, 1, {0, 0, 0}};
  for (int i = 0; i < tri(du); ++i) ia.sig_u[i] = h->sig_u_host[i];
  init_cells_kernel<<<nblocks((size_t)h->Bpad), 256, 0, h->stream>>>(h->recA, h->recB, ia, T - 1, 1, h->x0, h->sig_x0, dev, h->B);
  fill_kernel<<<nblocks((size_t)h->Bpad), 256, 0, h->stream>>>(h->alpha_cell + (size_t)slot * h->Bpad, (size_t)h->Bpad, alpha_init);
  h->launches += 2;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->cfg.z_per_problem) {
    FieldMap fz{dz, 0, 0, dz, 1, 0, 0, 1};
    rc = pack(h, h->z_cell, fz, T - 1, 1, z_new, false);
  } else {
    CUDA_OK(cudaMemcpyAsync(h->z_cell + (size_t)slot * dz, z_new, (size_t)dz * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  return rc;
}

static int ckf_step_impl(i2c_handle_t h, const double* y, const double* u, const double* sig_zeta, bool sync);
// I2C_B200_TRACE=1: host-side time of the phases of i2c_mpc_step (filter issue, sweeps issue, tail issue, wait), printed every
// 16 calls on stderr -- where a control step's time goes outside the kernels
struct StepTrace {
  static constexpr int N = 5;
  std::chrono::steady_clock::time_point t[N];
  bool on;
  StepTrace() : on(getenv("I2C_B200_TRACE") != nullptr) {}
  void mark(int i) {
    if (on) t[i] = std::chrono::steady_clock::now();
  }
  void done() {
    if (!on) return;
    static double acc[N] = {0, 0, 0, 0, 0};
    static int calls = 0;
    for (int i = 1; i < N; ++i) acc[i] += std::chrono::duration<double, std::micro>(t[i] - t[i - 1]).count();
    if (++calls % 16 == 0) {
      fprintf(stderr, "[i2c_mpc_step] us per call: filter issue %.1f, sweeps issue %.1f, tail issue %.1f, wait %.1f\n", acc[1] / 16,
              acc[2] / 16, acc[3] / 16, acc[4] / 16);
      for (double& a : acc) a = 0;
    }
  }
};
// One closed-loop MPC step with a single synchronisation: PartiallyObservedMpcPolicy.__call__
// (policy/mpc.py:156-182) = [filter] -> n_iter x (forward, backward, _update_priors) -> first action -> horizon shift.
int i2c_mpc_step(i2c_handle_t h, int32_t do_filter, const double* y, const double* u_prev, const double* sig_zeta,
                 int32_t n_iter, const double* z_new, const double* mu_u_init, double alpha_init, double* u_out) {
  REQUIRE(h && z_new && mu_u_init && u_out, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(!h->cfg.z_per_problem, "i2c_mpc_step expects shared cell targets (z_per_problem = 0)");
  const int dx = h->d.dx, du = h->d.du, n = dx + du, dz = h->d.dz, T = h->T;
  REQUIRE(dz <= 16 && du <= 16, "internal: by-value setter too small");
  int rc = 0;
  StepTrace tr;
  tr.mark(0);
  if (do_filter) {
    REQUIRE(y && u_prev && sig_zeta, "filter step needs y, u_prev, sig_zeta");
    rc = ckf_step_impl(h, y, u_prev, sig_zeta, false);  // stream-ordered; the single synchronisation is at the end
    if (rc) return rc;
  }
  tr.mark(1);
  const bool by_value = T <= 64;  // the cell flags ride in the tail kernel's arguments
  rc = run_core(h, n_iter, I2C_PH_FORWARD | I2C_PH_BACKWARD | I2C_PH_UPDATE_PRIORS, !by_value);
  if (rc) return rc;
  tr.mark(2);
  // first action = cells[0].mu_u0_m (policy/mpc.py:166) and the horizon shift (policy/mpc.py:174-181) in one launch: the popped
  // cell's ring slot IS the slot of the appended cell
  const double* latest = rec_latest(h);
  h->cell_head = (h->cell_head + 1) % T;
  const int slot = (T - 1 + h->cell_head) % T;
  h->flags[slot] = I2C_CELL_INDEPENDENT | I2C_CELL_EXPERT | I2C_CELL_OWN_ALPHA;
  h->index[slot] = 0;
  {
    SmallVals mu_v, z_v;
    SmallInts f_v;
    for (int i = 0; i < du; ++i) mu_v.v[i] = mu_u_init[i];
    for (int i = 0; i < dz; ++i) z_v.v[i] = z_new[i];
    for (int i = 0; i < T && i < 64; ++i) f_v.v[i] = h->flags[i];
    if (!by_value) CUDA_OK(cudaMemcpyAsync(h->cell_flags_dev, h->flags.data(), T * 4, cudaMemcpyHostToDevice, h->stream));
    InitArgs ia{dx, du, n, h->r.e_post(), T, h->cell_head, h->Bpad, h->ntiles, 1, {0, 0, 0}};
    for (int i = 0; i < tri(du); ++i) ia.sig_u[i] = h->sig_u_host[i];
    mpc_tail_kernel<<<nblocks((size_t)h->Bpad), 256, 0, h->stream>>>(h->recA, h->recB, latest, ia, h->x0, h->sig_x0, mu_v, h->B, slot,
                                                                    h->cell_flags_dev, h->cell_index_dev, f_v, by_value ? T : 0,
                                                                    h->alpha_cell, alpha_init, h->z_cell, z_v, dz, h->scratch);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(u_out, h->scratch, (size_t)h->B * du * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  h->launches += 1;
  tr.mark(3);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  tr.mark(4);
  tr.done();
  return 0;
}

static int ckf_step_impl(i2c_handle_t h, const double* y, const double* u, const double* sig_zeta, bool sync);
int i2c_ckf_step(i2c_handle_t h, const double* y, const double* u, const double* sig_zeta) {
  return ckf_step_impl(h, y, u, sig_zeta, true);
}
// sync = false: the caller synchronises the stream before the borrowed host buffers go out of scope (i2c_mpc_step)
static int ckf_step_impl(i2c_handle_t h, const double* y, const double* u, const double* sig_zeta, bool sync) {
  REQUIRE(h && y && u && sig_zeta, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(h->d.dy > 0, "this env defines no measurement map (only the quadrotor does)");
  const int dx = h->d.dx, du = h->d.du, dy = h->d.dy;
  // stage y and u in the scratch buffer in the caller's layout ([B][dy], [B][du]); the kernel reads them directly
  REQUIRE((size_t)h->B * (dy + du) <= h->scratch_elems, "internal: staging buffer too small");
  double* ty = h->scratch;
  double* tu = h->scratch + (size_t)h->B * dy;
  CUDA_OK(cudaMemcpyAsync(ty, y, (size_t)h->B * dy * 8, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(tu, u, (size_t)h->B * du * 8, cudaMemcpyHostToDevice, h->stream));
  CkfArgs a;
  memset(&a, 0, sizeof(a));
  a.x0 = h->x0;
  a.sig_x0 = h->sig_x0;
  a.y = ty;
  a.u = tu;
  a.canonical = 1;
  a.envpar = h->envpar;
  a.status = h->status;
  a.B = h->B;
  a.ntiles = h->ntiles;
  cubature_rule(1.0, 0.0, 0.0, dx, &a.sf, &a.w0, &a.wi);  // policy/mpc.py:121: CubatureQuadrature(1, 0, 0)
  for (int i = 0; i < tri(dx); ++i) a.sig_eta[i] = h->kp.sig_eta[i];
  full_to_tri(sig_zeta, dy, a.sig_zeta);
  int rc = launch_ckf(h->cfg.env, a, (void*)h->stream);
  h->launches += 1;
  if (rc != 0) return set_err(-100 - rc, "CKF kernel launch failed");
  if (sync) CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_gauss_hermite(int32_t degree, double* nodes, double* weights) {
  REQUIRE(degree >= 1 && degree <= MAX_GH && nodes && weights, "degree must be in [1, 8]");
  gauss_hermite_rule(degree, nodes, weights);
  return 0;
}

static int quadrature_impl(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S, double quad_alpha,
                           double quad_beta, double quad_kappa, int gh_degree, const double* env_par, double* m_y, double* S_y,
                           double* S_xy, int32_t* status, int32_t device);

int i2c_quadrature(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S, double quad_alpha,
                   double quad_beta, double quad_kappa, const double* env_par, double* m_y, double* S_y, double* S_xy,
                   int32_t* status, int32_t device) {
  return quadrature_impl(env, fn, n_problems, m, S, quad_alpha, quad_beta, quad_kappa, 0, env_par, m_y, S_y, S_xy, status, device);
}

int i2c_quadrature_gh(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S, int32_t degree,
                      const double* env_par, double* m_y, double* S_y, double* S_xy, int32_t* status, int32_t device) {
  REQUIRE(degree >= 1 && degree <= MAX_GH, "Gauss-Hermite degree must be in [1, 8]");
  return quadrature_impl(env, fn, n_problems, m, S, 1.0, 0.0, 0.0, degree, env_par, m_y, S_y, S_xy, status, device);
}

static int quadrature_impl(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S, double quad_alpha,
                           double quad_beta, double quad_kappa, int gh_degree, const double* env_par, double* m_y, double* S_y,
                           double* S_xy, int32_t* status, int32_t device) {
  REQUIRE(env >= 0 && env < I2C_ENV_COUNT, "unknown env id (no CPU fallback for unregistered callables)");
  REQUIRE(fn >= 0 && fn <= 3 && n_problems >= 1 && m && S && m_y && S_y && S_xy, "bad argument");
  const EnvDims& d = kEnv[env];
  REQUIRE(d.np == 0 || env_par, "this env needs env_par");
  REQUIRE(fn != 3 || d.dy > 0, "this env defines no measurement map");
  REQUIRE(fn != 1 || kHasTerm[env], "this env defines no terminal cost features");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(-2, "no CUDA device available: the i2c hot path has no CPU fallback");
  CUDA_OK(cudaSetDevice(device));
  const int n = d.dx + d.du;
  const int D = (fn == 0 || fn == 2) ? n : d.dx;
  const int DY = fn == 0 ? d.dz : (fn == 1 ? d.dzt : (fn == 2 ? d.dx : d.dy));
  const int B = n_problems, Bpad = (int)align_up((size_t)B, TILE), nt = Bpad / TILE;
  const int np = d.np > 0 ? d.np : 1;
  size_t canon = (size_t)B * (D + D * D + DY + DY * DY + D * DY + np);
  size_t tiled = (size_t)Bpad * (D + tri(D) + DY + tri(DY) + D * DY + np);
  double* buf = nullptr;
  int32_t* st = nullptr;
  CUDA_OK(cudaMalloc((void**)&buf, (canon + tiled) * 8));
  CUDA_OK(cudaMalloc((void**)&st, (size_t)Bpad * 4));
  double *c_m = buf, *c_S = c_m + (size_t)B * D, *c_my = c_S + (size_t)B * D * D, *c_Sy = c_my + (size_t)B * DY,
         *c_Sxy = c_Sy + (size_t)B * DY * DY, *c_par = c_Sxy + (size_t)B * D * DY;
  double *t_m = buf + canon, *t_S = t_m + (size_t)Bpad * D, *t_my = t_S + (size_t)Bpad * tri(D),
         *t_Sy = t_my + (size_t)Bpad * DY, *t_Sxy = t_Sy + (size_t)Bpad * tri(DY), *t_par = t_Sxy + (size_t)Bpad * D * DY;
  cudaStream_t s = 0;
  int rc = 0;
  do {
    if (cudaMemcpyAsync(c_m, m, (size_t)B * D * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(c_S, S, (size_t)B * D * D * 8, cudaMemcpyHostToDevice, s) != cudaSuccess) {
      rc = set_err(-3, "H2D copy failed");
      break;
    }
    FieldMap fm{D, 0, 0, D, 1, 0, 0, 0}, fS{tri(D), 0, 1, D, D, 0, 0, 0};
    pack_kernel<<<nblocks((size_t)Bpad * D), 256, 0, s>>>(t_m, fm, 0, 1, 1, 0, B, Bpad, nt, c_m, 0);
    pack_kernel<<<nblocks((size_t)Bpad * D * D), 256, 0, s>>>(t_S, fS, 0, 1, 1, 0, B, Bpad, nt, c_S, 0);
    if (d.np > 0) {
      cudaMemcpyAsync(c_par, env_par, (size_t)B * np * 8, cudaMemcpyHostToDevice, s);
      FieldMap fp{np, 0, 0, np, 1, 0, 0, 0};
      pack_kernel<<<nblocks((size_t)Bpad * np), 256, 0, s>>>(t_par, fp, 0, 1, 1, 0, B, Bpad, nt, c_par, 0);
    }
    QuadArgs a;
    a.m = t_m, a.S = t_S, a.envpar = t_par, a.my = t_my, a.Sy = t_Sy, a.Sxy = t_Sxy, a.status = st;
    a.B = B, a.ntiles = nt;
    cubature_rule(quad_alpha, quad_beta, quad_kappa, D, &a.sf, &a.w0, &a.wi);
    fill_gh(&a.gh, gh_degree);
    int lrc = launch_quadrature(env, fn, a, (void*)s);
    if (lrc != 0) {
      rc = set_err(-100 - lrc, "quadrature kernel launch failed");
      break;
    }
    FieldMap fy{DY, 0, 0, DY, 1, 0, 0, 0}, fSy{tri(DY), 0, 1, DY, DY, 0, 0, 0}, fxy{D * DY, 0, 0, D, DY, 0, 0, 0};
    unpack_kernel<<<nblocks((size_t)B * DY), 256, 0, s>>>(t_my, fy, 0, 1, 1, 0, B, nt, c_my);
    unpack_kernel<<<nblocks((size_t)B * DY * DY), 256, 0, s>>>(t_Sy, fSy, 0, 1, 1, 0, B, nt, c_Sy);
    unpack_kernel<<<nblocks((size_t)B * D * DY), 256, 0, s>>>(t_Sxy, fxy, 0, 1, 1, 0, B, nt, c_Sxy);
    cudaMemcpyAsync(m_y, c_my, (size_t)B * DY * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(S_y, c_Sy, (size_t)B * DY * DY * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(S_xy, c_Sxy, (size_t)B * D * DY * 8, cudaMemcpyDeviceToHost, s);
    if (status) cudaMemcpyAsync(status, st, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = set_err(-100 - (int)e, std::string("quadrature: ") + cudaGetErrorString(e));
  } while (0);
  cudaFree(buf);
  cudaFree(st);
  return rc;
}

// ---- snapshot / restore: [host state blob][workspace bytes]
// environment, inference kind, ABI version and batch size folded into one word of the snapshot header
static int32_t snap_tag(i2c_handle_t h) {
  return (int32_t)((((uint32_t)I2C_ABI_VERSION & 0xf) << 28) ^ (((uint32_t)h->cfg.env & 0xf) << 24) ^
                   (((uint32_t)h->cfg.inference & 0xf) << 20) ^ ((uint32_t)h->B & 0xfffff));
}

struct SnapHeader {
  uint64_t magic, ws_bytes;
  int32_t prior_is_A, latest_is_A, cell_head, tau, last_n_iter, problem_set, T, pad;
  double temp, dtemp;
  KParams kp;
};

int i2c_snapshot_bytes(i2c_handle_t h, size_t* bytes) {
  REQUIRE(h && bytes, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  *bytes = sizeof(SnapHeader) + (size_t)h->T * 8 + 3 * 8 + h->ws_bytes;
  return 0;
}

int i2c_snapshot(i2c_handle_t h, void* host_buf, size_t bytes) {
  REQUIRE(h && host_buf, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  size_t need;
  i2c_snapshot_bytes(h, &need);
  REQUIRE(bytes >= need, "snapshot buffer too small");
  if (int rc = normalise_belief(h)) return rc;
  SnapHeader hd;
  memset(&hd, 0, sizeof(hd));
  hd.magic = 0x6932635f62323030ull;
  hd.ws_bytes = h->ws_bytes;
  hd.prior_is_A = h->prior_is_A, hd.latest_is_A = h->latest_is_A, hd.cell_head = h->cell_head, hd.tau = h->tau;
  hd.last_n_iter = h->last_n_iter, hd.problem_set = h->problem_set, hd.T = h->T;
  hd.temp = h->temp, hd.dtemp = h->dtemp;
  hd.pad = snap_tag(h);
  hd.kp = h->kp;
  char* p = (char*)host_buf;
  memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  memcpy(p, h->flags.data(), (size_t)h->T * 4);
  p += (size_t)h->T * 4;
  memcpy(p, h->index.data(), (size_t)h->T * 4);
  p += (size_t)h->T * 4;
  double su[3] = {0, 0, 0};
  for (size_t i = 0; i < h->sig_u_host.size() && i < 3; ++i) su[i] = h->sig_u_host[i];
  memcpy(p, su, 24);
  p += 24;
  CUDA_OK(cudaMemcpyAsync(p, h->ws, h->ws_bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_restore(i2c_handle_t h, const void* host_buf, size_t bytes) {
  REQUIRE(h && host_buf, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  REQUIRE(bytes >= sizeof(SnapHeader), "snapshot too small");
  SnapHeader hd;
  const char* p = (const char*)host_buf;
  memcpy(&hd, p, sizeof(hd));
  REQUIRE(hd.magic == 0x6932635f62323030ull && hd.ws_bytes == h->ws_bytes && hd.T == h->T,
          "snapshot does not match this handle's configuration");
  REQUIRE(hd.pad == snap_tag(h), "snapshot was taken from a handle with another environment / inference / batch / ABI");
  REQUIRE(bytes >= sizeof(SnapHeader) + (size_t)h->T * 8 + 3 * 8 + h->ws_bytes, "snapshot buffer is truncated");
  REQUIRE(tri(h->d.du) <= 3, "internal: sig_u blob holds tri(du) <= 3 doubles");
  p += sizeof(hd);
  h->prior_is_A = hd.prior_is_A, h->latest_is_A = hd.latest_is_A, h->cell_head = hd.cell_head, h->tau = hd.tau;
  h->last_n_iter = hd.last_n_iter, h->problem_set = hd.problem_set != 0;
  h->temp = hd.temp, h->dtemp = hd.dtemp;
  h->kp = hd.kp;
  h->flags.resize(h->T);
  h->index.resize(h->T);
  memcpy(h->flags.data(), p, (size_t)h->T * 4);
  p += (size_t)h->T * 4;
  memcpy(h->index.data(), p, (size_t)h->T * 4);
  p += (size_t)h->T * 4;
  double su[3];
  memcpy(su, p, 24);
  p += 24;
  h->sig_u_host.assign(su, su + tri(h->d.du));
  // the snapshot holds the belief in the primary buffers (normalise_belief): point this handle at them
  CUDA_OK(cudaStreamSynchronize(h->up_stream));
  if (h->belief_swapped) {
    std::swap(h->x0, h->x0_alt);
    std::swap(h->sig_x0, h->sig_x0_alt);
    h->belief_swapped = false;
  }
  CUDA_OK(cudaMemcpyAsync(h->ws, p, h->ws_bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int i2c_rollout(int32_t env, int32_t n_problems, int32_t n_rollouts, int32_t horizon, const double* x_init,
                const double* K, const double* k, const double* sigK, const double* expert_mu, const double* expert_lam,
                int32_t soft_expert, const double* eta, const double* eps_u, const double* sig_eta, uint64_t seed,
                const double* env_par, double* xu, double* z, double* z_term, double* x_final, int32_t device) {
  REQUIRE(env >= 0 && env < I2C_ENV_COUNT, "unknown env id (no CPU fallback for unregistered envs)");
  REQUIRE(n_problems >= 1 && n_rollouts >= 1 && horizon >= 1 && x_init && K && k && xu && z, "bad argument");
  REQUIRE((expert_mu == nullptr) == (expert_lam == nullptr), "expert policy needs both mu and lam");
  const EnvDims& d = kEnv[env];
  REQUIRE(d.np == 0 || env_par, "this env needs env_par");
  REQUIRE(eta || sig_eta, "give either the disturbances eta or sig_eta for the device RNG (zero matrix = noise free)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(-2, "no CUDA device available: the i2c hot path has no CPU fallback");
  CUDA_OK(cudaSetDevice(device));
  const size_t B = n_problems, R = n_rollouts, T = horizon, dx = d.dx, du = d.du, n = dx + du, dz = d.dz, dzt = d.dzt;
  const size_t np = d.np > 0 ? d.np : 0;
  RolloutArgs a;
  memset(&a, 0, sizeof(a));
  a.B = n_problems, a.R = n_rollouts, a.T = horizon, a.soft_expert = soft_expert, a.seed = seed;
  a.hard_threshold = 3.0;  // ExpertTimeIndexedLinearGaussianPolicy.hard_exp_threshold (policy/linear.py:48)
  if (!eta) {
    std::vector<double> L(sig_eta, sig_eta + dx * dx);
    bool zero = true;
    for (double v : L) zero = zero && v == 0.0;
    a.noise_free = zero;
    if (!zero) {
      REQUIRE(host_chol(L, (int)dx), "sig_eta must be positive definite");
      for (size_t i = 0; i < dx; ++i)
        for (size_t j = 0; j <= i; ++j) a.chol_eta[i * (i + 1) / 2 + j] = L[i * dx + j];
    }
  }
  struct Buf { const double* host; size_t elems; const double** dev; };
  const size_t sizes_in[9] = {B * R * dx, B * T * du * dx, B * T * du, sigK ? B * T * du * du : 0, expert_mu ? B * T * dx : 0,
                              expert_lam ? B * T * dx * dx : 0, eta ? B * R * T * dx : 0, eps_u ? B * R * T * du : 0, B * np};
  const double* hosts[9] = {x_init, K, k, sigK, expert_mu, expert_lam, eta, eps_u, env_par};
  const double** devs[9] = {&a.x_init, &a.K, &a.k, &a.sigK, &a.ex_mu, &a.ex_lam, &a.eta, &a.eps_u, &a.envpar};
  const size_t n_xu = B * R * T * n, n_z = B * R * T * dz, n_zt = B * R * dzt, n_xf = B * R * dx;
  size_t total = n_xu + n_z + n_zt + n_xf;
  for (size_t v : sizes_in) total += v;
  double* buf = nullptr;
  CUDA_OK(cudaMalloc((void**)&buf, total * 8));
  size_t off = 0;
  int rc = 0;
  for (int i = 0; i < 9 && rc == 0; ++i) {
    if (sizes_in[i] == 0) continue;
    *devs[i] = buf + off;
    if (cudaMemcpy(buf + off, hosts[i], sizes_in[i] * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = set_err(-3, "H2D copy failed");
    off += sizes_in[i];
  }
  a.xu = buf + off, a.z = a.xu + n_xu, a.z_term = a.z + n_z, a.x_final = a.z_term + n_zt;
  if (rc == 0) {
    int lrc = launch_rollout(env, a, nullptr);
    if (lrc != 0) rc = set_err(-100 - lrc, "rollout kernel launch failed");
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(xu, a.xu, n_xu * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(z, a.z, n_z * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && z_term && kHasTerm[env]) e = cudaMemcpy(z_term, a.z_term, n_zt * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && x_final) e = cudaMemcpy(x_final, a.x_final, n_xf * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = set_err(-100 - (int)e, std::string("rollout: ") + cudaGetErrorString(e));
  }
  cudaFree(buf);
  return rc;
}

// Device math primitives of csrc/fastmath.cuh evaluated on host-provided arguments (accuracy tests: tests/test_gpu_fastmath.py)
__global__ void fastmath_probe_kernel(int fn, int n, const double* x, double* y0, double* y1) {
  using namespace i2c;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0, b = 0.0;
  switch (fn) {
    case 0: a = fast_rsqrt(x[i]); break;
    case 1: a = fast_rcp(x[i]); break;
    case 2: a = fast_exp_neg(x[i]); break;
    case 3: a = fast_exp_neg_lat(x[i]); break;
    case 4: fast_sincos(x[i], &a, &b); break;
    case 5: seq_sincos(x[i], &a, &b); break;
    case 6: {  // log-determinant accumulator: log of the running product of x[i], x[i]^2
      LogAcc l;
      l.reset();
      l.mul(x[i]);
      l.mul(x[i] * x[i]);
      a = l.value();
      break;
    }
    default: break;
  }
  y0[i] = a;
  y1[i] = b;
}

int i2c_fastmath_probe(int32_t device, int32_t fn, int32_t n, const double* x, double* y0, double* y1) {
  REQUIRE(x && y0 && y1 && n > 0 && fn >= 0 && fn <= 6, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(-2, "no CUDA device available");
  int prev = 0;
  cudaGetDevice(&prev);
  CUDA_OK(cudaSetDevice(device));
  double* d = nullptr;
  CUDA_OK(cudaMalloc((void**)&d, (size_t)n * 8 * 3));
  CUDA_OK(cudaMemcpy(d, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  fastmath_probe_kernel<<<(n + 127) / 128, 128>>>(fn, n, d, d + n, d + 2 * (size_t)n);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(y0, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(y1, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaSetDevice(prev);
  CUDA_OK(e);
  return 0;
}

int i2c_dfma_peak(int32_t device, double* tflops) {
  REQUIRE(tflops != nullptr, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(-2, "no CUDA device available");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
  double* out = nullptr;
  CUDA_OK(cudaMalloc((void**)&out, (size_t)threads * blocks * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, 0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  CUDA_OK(cudaGetLastError());
  *tflops = best;
  return 0;
}

int i2c_host_alloc(size_t bytes, void** out) {
  REQUIRE(out && bytes > 0, "bad argument");
  CUDA_OK(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return 0;
}

int i2c_host_free(void* p) {
  if (p) CUDA_OK(cudaFreeHost(p));
  return 0;
}

int i2c_kernel_launches(i2c_handle_t h, int64_t* n) {
  REQUIRE(h && n, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  *n = h->launches;
  return 0;
}

int i2c_last_run_ms(i2c_handle_t h, float* ms) {
  REQUIRE(h && ms, "NULL argument");
  DeviceGuard device_guard_(h->cfg.device);
  CUDA_OK(cudaEventSynchronize(h->ev1));
  CUDA_OK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}

}  // extern "C"
