/*
 * i2c_b200.h -- C-ABI of the B200-native Gaussian i2c EM sweep.
 *
 * The reference (JoeMWatson/input-inference-for-control) is pure Python/NumPy and has no
 * FFI/plugin boundary of its own; the hot path sits behind ordinary Python classes.  This header
 * is the NEW boundary directly underneath that Python surface: every entry point cites the
 * reference interface it replaces (paths relative to the reference root).  All pointers are plain
 * host pointers to C-contiguous fp64 / int32 arrays unless the name ends in `_dev`; no torch /
 * C++ types cross the boundary.  INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - every call returns 0 on success, <0 on API misuse or CUDA error (text: i2c_last_error());
 *   - NUMERICAL failures (non-PD Cholesky, singular pdf ratio, NaN alpha, det<=0) are reported
 *     per problem through i2c_get_status(), never as a call error: one bad problem does not
 *     poison its neighbours (the reference raises LinAlgError / ValueError instead:
 *     i2c/inference/quadrature.py:17-24, i2c/i2c.py:950-951,1074-1077);
 *   - a handle is bound to one CUDA device, is not thread-safe, and all work is ordered on the
 *     stream given at creation; calls that return data to the host synchronise that stream;
 *   - host arrays are "canonical" layout: problem-major, then cell, then row-major matrix,
 *     e.g. K[B][T][du][dx]; the library converts to / from its tiled device layout itself.
 */
#ifndef I2C_B200_H
#define I2C_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2C_ABI_VERSION 1

/* Environment registry: i2c/model.py:25-36 (make_env_model keys) + scripts/mpc_state_est/mpc_quad.py:219-386 */
enum i2c_env {
  I2C_ENV_LINEAR = 0,            /* "LinearKnown"               env_def.py:139-191, model.py:226-242 */
  I2C_ENV_LINEAR_MIN_ENERGY = 1, /* "LinearKnownMinimumEnergy"  env_def.py:194-230 */
  I2C_ENV_PENDULUM = 2,          /* "PendulumKnown"             env_def.py:233-309, env_autograd.py:5-19 */
  I2C_ENV_PENDULUM_ACT_REG = 3,  /* "PendulumKnownActReg"       env_def.py:312-346 */
  I2C_ENV_CARTPOLE = 4,          /* "CartpoleKnown"             env_def.py:491-612, env_autograd.py:25-54 */
  I2C_ENV_DOUBLE_CARTPOLE = 5,   /* "DoubleCartpoleKnown"       env_def.py:615-761, env_autograd.py:60-167 */
  I2C_ENV_QUADROTOR = 6,         /* QuadrotorKnown (fp64 restatement of the Box2D step) mpc_quad.py:219-386 */
  I2C_ENV_COUNT = 7
};

/* Inference kinds: i2c/exp_types.py:25-49 */
/* I2C_INF_GAUSS_HERMITE: GaussHermiteQuadrature(degree) (exp_types.py:52-68), degree passed in i2c_config.quad_alpha */
enum i2c_inference { I2C_INF_CUBATURE = 0, I2C_INF_LINEARIZE = 1, I2C_INF_GAUSS_HERMITE = 2 };

/* Per-problem status words (replace the reference's exceptions) */
enum i2c_status {
  I2C_OK = 0,
  I2C_FAIL_CHOL_PRIOR = 1,    /* joint prior not PD            quadrature.py:17-24 via i2c.py:391 */
  I2C_FAIL_CHOL_OBS = 2,      /* S_z + Sigma_xi not PD         i2c.py:398 */
  I2C_FAIL_CHOL_FILTERED = 3, /* Sigma_xu1_f not PD            quadrature.py:17-24 via i2c.py:415 */
  I2C_FAIL_CHOL_X3 = 4,       /* Sigma_x3_f not PD             i2c.py:423 */
  I2C_FAIL_CHOL_TERMINAL = 5, /* terminal update               i2c.py:432-438 */
  I2C_FAIL_CHOL_POSTERIOR = 6,/* Sigma_xu1_m not PD            quadrature.py:17-24 via i2c.py:594 */
  I2C_FAIL_MVN = 7,           /* pdf-ratio covariance singular i2c.py:369-374 (scipy multivariate_normal) */
  I2C_FAIL_NAN_ALPHA = 8,     /* alpha update is NaN           i2c.py:950-951 */
  I2C_FAIL_POLICY_DET = 9,    /* det(Sigma_u0_m) <= 0          i2c.py:1074-1077 */
  I2C_FAIL_CHOL_PROPAGATE = 10, /* propagate joint not PD      i2c.py:181-197 */
  I2C_FAIL_COV_CONTROL = 11,  /* covariance-control terminal solve  i2c.py:548-559 */
  I2C_FAIL_CKF = 12           /* cubature Kalman filter        policy/mpc.py:125-145 */
};

/* Phases of one EM iteration; i2c_run executes the selected ones in this order. */
enum i2c_phase {
  I2C_PH_FORWARD = 1,        /* I2cGraph._forward_msgs        i2c.py:876-880 (cells: :350-447) */
  I2C_PH_BACKWARD = 2,       /* I2cGraph._backward_msgs       i2c.py:882-886 (cells: :544-610) */
  I2C_PH_PROPAGATE = 4,      /* I2cGraph.propagate            i2c.py:1247-1251 (cells: :150-199) */
  I2C_PH_MSTEP = 8,          /* calc_cost + compute_update_alpha(update) + metrics  i2c.py:1004-1027 */
  I2C_PH_UPDATE_PRIORS = 16, /* I2cGraph._update_priors       i2c.py:1210-1221 */
  I2C_PH_CALIBRATE = 32,     /* calibrate_alpha (after PROPAGATE)  i2c.py:895-911 */
  I2C_PH_ONLY_DECREASE = 64, /* calibrate_alpha(only_decrease=True) */
  I2C_PH_STORE_AUX = 128,    /* also write the per-cell auxiliary messages (mu_z0_f, sig_z0_f, prior joint, ...) */
  I2C_PH_RICCATI = 256       /* I2cGraph._backward_ricatti_msgs (i2c.py:888-893, cells :612-678); Linearize, linear envs,
                                needs the aux records of the preceding forward/backward pass */
};
/* I2cGraph.learn_msgs (i2c.py:1238-1245) without / with propagate */
#define I2C_PH_LEARN (I2C_PH_FORWARD | I2C_PH_BACKWARD | I2C_PH_MSTEP | I2C_PH_UPDATE_PRIORS)
#define I2C_PH_LEARN_PROPAGATE (I2C_PH_LEARN | I2C_PH_PROPAGATE)

/* Per-cell flags (the reference keeps these as attributes on each I2cCell: i2c.py:82,132,143) */
enum i2c_cell_flag {
  I2C_CELL_INDEPENDENT = 1, /* state_action_independence */
  I2C_CELL_TERMINAL = 2,    /* terminal_cell */
  I2C_CELL_EXPERT = 4,      /* use_expert_controller (propagate only; the quadrature forward pass ignores it) */
  I2C_CELL_OWN_ALPHA = 8    /* cell carries its own sig_xi scale (MPC horizon shift, policy/mpc.py:174-176) */
};

/* Per-cell device fields readable / writable in canonical layout (i2c_get_field / i2c_set_field).
 * "tri" covariances are returned as full symmetric [d][d] matrices. */
enum i2c_field {
  /* posterior record written by the backward sweep */
  I2C_F_MU_XU0_M = 0, I2C_F_SIG_XU0_M = 1, I2C_F_K = 2, I2C_F_KK = 3, I2C_F_SIGK = 4,
  /* prior record read by the forward sweep (mu_xu0_f / sig_xu0_f after _update_priors, and the K it uses) */
  I2C_F_PRIOR_MU = 5, I2C_F_PRIOR_SIG = 6, I2C_F_PRIOR_K = 7,
  /* filtered record written by the forward sweep */
  I2C_F_MU_XU1_F = 8, I2C_F_SIG_XU1_F = 9, I2C_F_MU_X3_F = 10, I2C_F_SIG_X3_F = 11, I2C_F_J_DYN = 12,
  /* auxiliary messages (only valid after a run with I2C_PH_STORE_AUX) */
  I2C_F_MU_XU0_F = 13, I2C_F_SIG_XU0_F = 14, I2C_F_MU_Z0_F = 15, I2C_F_SIG_Z0_F = 16,
  I2C_F_MU_Z0_M = 17, I2C_F_SIG_Z0_M = 18, I2C_F_MU_X3_M = 19, I2C_F_SIG_X3_M = 20,
  /* propagate record (after a run with I2C_PH_PROPAGATE | I2C_PH_STORE_AUX) */
  I2C_F_MU_XU0_PF = 21, I2C_F_SIG_XU0_PF = 22, I2C_F_MU_Z0_PF = 23, I2C_F_SIG_Z0_PF = 24,
  I2C_F_MU_X3_PF = 25, I2C_F_SIG_X3_PF = 26,
  /* terminal cost-feature moments of the last cell: shape [B][1][...] */
  I2C_F_MU_Z3_M = 27, I2C_F_SIG_Z3_M = 28,
  /* Riccati messages (after I2C_PH_RICCATI): full [dx][dx] matrices / [dx] vectors */
  I2C_F_LAMBDA_X3_B = 29, I2C_F_NU_X3_B = 30, I2C_F_LAMBDA_X0_B = 31, I2C_F_NU_X0_B = 32,
  I2C_F_COUNT = 33
};

/* Per-iteration, per-problem scalars (the Python lists on I2cGraph: i2c.py:1329-1372) */
enum i2c_metric {
  I2C_M_ALPHA = 0, I2C_M_ALPHA_DESIRED = 1, I2C_M_ALPHA_PF = 2, I2C_M_COST_M = 3, I2C_M_COST_M_VAR = 4,
  I2C_M_COST_PF = 5, I2C_M_COST_PF_VAR = 6, I2C_M_COST_PF_MIN = 7, I2C_M_POLICY_ENTROPY = 8,
  I2C_M_X_PRIOR_ENTROPY = 9, I2C_M_PROPAGATE_ENTROPY = 10, I2C_M_KL_TERM = 11, I2C_M_COUNT = 12
};

typedef struct i2c_handle_s* i2c_handle_t;

/* Static configuration: the arguments of I2cGraph.__init__ (i2c.py:735-750) that fix shapes. */
typedef struct i2c_config {
  int32_t abi_version; /* I2C_ABI_VERSION */
  int32_t env;         /* enum i2c_env */
  int32_t inference;   /* enum i2c_inference */
  int32_t n_problems;  /* B: independent problems (1 for the reference's single-trajectory use) */
  int32_t horizon;     /* H */
  int32_t max_iters;   /* capacity of the per-iteration metric ring written by one i2c_run */
  int32_t device;      /* CUDA device ordinal */
  int32_t z_per_problem; /* 0: cell targets z[t] shared by all problems; 1: z[b][t] */
  int32_t enable_aux;    /* 1: allocate the auxiliary / propagate message records (I2C_PH_STORE_AUX usable) */
  double quad_alpha, quad_beta, quad_kappa; /* CubatureQuadrature(alpha, beta, kappa) exp_types.py:30-49 */
} i2c_config_t;

/* Dimensions of an environment (dim_x, dim_u, dim_z, dim_z_term, n env parameters per problem, dim_y). */
int i2c_env_dims(int32_t env, int32_t* dx, int32_t* du, int32_t* dz, int32_t* dzt, int32_t* n_par, int32_t* dy);

/* Bytes of device workspace a handle needs (so the caller can own the allocation, e.g. a torch tensor). */
int i2c_workspace_bytes(const i2c_config_t* cfg, size_t* bytes);

/* Create a handle.  workspace_dev == NULL: the library cudaMallocs its own workspace.
 * stream: a cudaStream_t (CUstream) passed as void*; NULL = the device's default stream. */
int i2c_create(const i2c_config_t* cfg, void* workspace_dev, size_t workspace_bytes, void* stream, i2c_handle_t* out);
int i2c_destroy(i2c_handle_t h);

/* Problem definition == remaining arguments of I2cGraph.__init__ (i2c.py:735-750) plus the model
 * constants it reads from `sys` (x0, sig_x0, sig_eta, zg, zg_term).  Resets every cell to its
 * constructor state (I2cCell.__init__, i2c.py:54-148): priors = (mu_u, sig_u), K = 0, all cells
 * independent, last cell terminal, tau = H-1, temp = 1.
 *   x0[B][dx], sig_x0[B][dx][dx], sig_eta[dx][dx], mu_u[B][H][du], sig_u[du][du],
 *   QR[dz][dz] (= block_diag(Q,R) or R), Qf[dzt][dzt] or NULL, z[H][dz] (or [B][H][dz]), z_graph[dz] (the graph-level
 *   target used by calc_cost), z_term[dzt] (or [B][dzt] when z_per_problem) or NULL,
 *   alpha0[B], mu_x_term[dx]/sig_x_term[dx][dx] or NULL (covariance control),
 *   env_par[B][n_par] or NULL (linear envs: A row-major, B, a). */
int i2c_set_problem(i2c_handle_t h, const double* x0, const double* sig_x0, const double* sig_eta,
                    const double* mu_u, const double* sig_u, const double* QR, const double* Qf,
                    const double* z, const double* z_graph, const double* z_term, const double* alpha0,
                    double alpha_update_tol, const double* mu_x_term, const double* sig_x_term, double dtemp,
                    const double* env_par);

/* sys.x0 / sys.sig_x0 are re-read at the start of every sweep (i2c.py:877-878); MPC overwrites them
 * (policy/mpc.py:149-150). */
int i2c_set_initial_state(i2c_handle_t h, const double* x0, const double* sig_x0);
/* Same without the trailing synchronisation: the (pinned) host buffers must stay valid and unchanged until the next
 * synchronising call on the handle (i2c_synchronize or any getter); stream-ordered before the next i2c_run. */
int i2c_set_initial_state_async(i2c_handle_t h, const double* x0, const double* sig_x0);
int i2c_set_initial_state_dev(i2c_handle_t h, const double* x0_dev, const double* sig_x0_dev);

/* Per-cell flags / indices / tau (attributes the scripts set on cells and graph). */
int i2c_set_cell_flags(i2c_handle_t h, const int32_t* flags /*[H]*/);
int i2c_get_cell_flags(i2c_handle_t h, int32_t* flags /*[H]*/);
int i2c_set_cell_index(i2c_handle_t h, const int32_t* index /*[H]*/);
int i2c_set_tau(i2c_handle_t h, int32_t tau);
/* Per-cell targets: `c.z = z_traj[i]` (policy/mpc.py:30-31).  z[H][dz] (or [B][H][dz] when z_per_problem). */
int i2c_set_cell_targets(i2c_handle_t h, const double* z);
int i2c_set_alpha(i2c_handle_t h, const double* alpha /*[B]*/); /* _override_alpha / update_xi on all cells */
int i2c_get_alpha(i2c_handle_t h, double* alpha /*[B]*/);
int i2c_set_temp(i2c_handle_t h, double temp);
int i2c_get_temp(i2c_handle_t h, double* temp);

/* Run n_iter iterations of the selected phases in ONE persistent kernel launch (all sweeps over the
 * horizon and all iterations stay on the device).  Replaces the Python loops of
 * I2cGraph.learn_msgs / _forward_backward_msgs / propagate / _maximize / calibrate_alpha and
 * MpcPolicy.optimize (policy/mpc.py:147-154).  Asynchronous on the handle's stream. */
int i2c_run(i2c_handle_t h, int32_t n_iter, int32_t phases);

/* Parallel-in-time variant of i2c_run for long horizons (not in the reference, whose sweeps are sequential loops over the
 * cells, i2c.py:876-886): the horizon is cut into chunks of `chunk_cells` cells; per sweep (1) every chunk composes the
 * Kalman filtering / RTS smoothing elements of its cells (associative operators), (2) the boundary messages are pushed
 * through the chunk aggregates, (3) every chunk runs the ordinary cell recursion from its exact boundary message, so the
 * records, metrics and alpha schedule are those of i2c_run up to round-off.  Sequential depth 2*chunk_cells + H/chunk_cells
 * cells instead of H.  Exact only where the cell map is linear-Gaussian in the state message: I2C_INF_LINEARIZE on
 * I2C_ENV_LINEAR / I2C_ENV_LINEAR_MIN_ENERGY with cells that are independent or non-expert feedback cells; anything else
 * is refused (no fallback).  phases: I2C_PH_FORWARD | I2C_PH_BACKWARD [| I2C_PH_MSTEP | I2C_PH_UPDATE_PRIORS | I2C_PH_STORE_AUX]. */
int i2c_run_scan(i2c_handle_t h, int32_t n_iter, int32_t phases, int32_t chunk_cells);

/* Wait for the stream; number of metric rows written by the last i2c_run. */
int i2c_synchronize(i2c_handle_t h);

/* Per-iteration scalars of the last i2c_run: out[n_iter][B]. */
int i2c_get_metric(i2c_handle_t h, int32_t metric, double* out, int32_t n_iter);
/* Several metrics with ONE synchronisation: out[n_metrics][n_iter][B] (e.g. cost and alpha of the last iteration, the
 * per-step read-back of an EM loop: i2c_run.py:109-113 reads i2c.costs_m[-1], i2c.alphas[-1] every iteration). */
int i2c_get_metrics(i2c_handle_t h, const int32_t* metrics, int32_t n_metrics, double* out, int32_t n_iter);
/* Pipelined variant for host loops that log a step's cost / alpha (scripts/i2c_run.py:84-88 prints them per iteration) but do
 * not feed them back: gathers the metrics of the LAST iteration of the most recent i2c_run into staging slot `slot` (0 / 1) and
 * copies them to the page-locked `out` [n_metrics][B] on the copy stream.  i2c_metrics_wait(slot) blocks until that copy has
 * landed.  With two slots the host queues step i+1 before it collects step i: the stream never drains. */
int i2c_get_last_metrics_async(i2c_handle_t h, const int32_t* metrics, int32_t n_metrics, double* out, int32_t slot);
int i2c_metrics_wait(i2c_handle_t h, int32_t slot);
int i2c_get_status(i2c_handle_t h, int32_t* status /*[B]*/, int32_t* info /*[B]: (iter<<16 | cell)*/);
/* The per-problem status / info words are sticky (the kernels only write them while they are OK): clear them, e.g. after the
 * caller handled a failure (the reference raises LinAlgError once, quadrature.py:17-24; a later sweep starts clean). */
int i2c_clear_status(i2c_handle_t h);

/* Cell-attribute views (the attributes scripts read off `i2c.cells[t]`), cells [t0, t1). */
int i2c_get_field(i2c_handle_t h, int32_t field, int32_t t0, int32_t t1, double* out);
int i2c_set_field(i2c_handle_t h, int32_t field, int32_t t0, int32_t t1, const double* in);
int i2c_field_shape(i2c_handle_t h, int32_t field, int32_t* rows, int32_t* cols);

/* Controller extraction == I2cGraph.get_local_linear_policy (i2c.py:1253-1264):
 * K[B][H][du][dx], k[B][H][du], sigK[B][H][du][du]; any pointer may be NULL. */
int i2c_get_policy(i2c_handle_t h, double* K, double* k, double* sigK);
/* Asynchronous variant: the controllers of the latest sweep are converted on the compute stream and copied to
 * (ideally pinned) host arrays on a separate copy stream, so the transfer overlaps the next i2c_run; the arrays are
 * valid after i2c_copy_wait().  One transfer may be in flight per handle. */
int i2c_get_policy_async(i2c_handle_t h, double* K, double* k, double* sigK);
int i2c_copy_wait(i2c_handle_t h);
/* Same, into DEVICE buffers in canonical layout (for the NCCL gather of controllers). */
int i2c_get_policy_dev(i2c_handle_t h, double* K_dev, double* k_dev, double* sigK_dev);

/* MPC horizon shift == cells.pop(0); cells.append(deepcopy(cell_init)) with the new cell's target
 * (policy/mpc.py:174-181).  The appended cell is the constructor-state cell (initial mu_u/sig_u prior,
 * K = 0, independent, not terminal, index 0, alpha = alpha_init).  z_new[dz] (or [B][dz]). */
int i2c_shift_horizon(i2c_handle_t h, const double* z_new, const double* mu_u_init /*[du]*/, double alpha_init);

/* Cubature Kalman filter step == PartiallyObservedMpcPolicy.filter (policy/mpc.py:125-145) on the
 * handle's belief (x0, sig_x0): predict through the dynamics with u fixed, update on measure(x).
 * y[B][dy], u[B][du], sig_zeta[dy][dy]. */
int i2c_ckf_step(i2c_handle_t h, const double* y, const double* u, const double* sig_zeta);
/* One whole control step of PartiallyObservedMpcPolicy.__call__ (policy/mpc.py:156-182) with a single host
 * synchronisation: [do_filter: i2c_ckf_step(y, u_prev)] -> n_iter x (forward, backward, _update_priors) ->
 * u_out[B][du] = cells[0].mu_u0_m -> horizon shift with the new last cell's target z_new[dz]. */
int i2c_mpc_step(i2c_handle_t h, int32_t do_filter, const double* y, const double* u_prev, const double* sig_zeta,
                 int32_t n_iter, const double* z_new, const double* mu_u_init, double alpha_init, double* u_out);
int i2c_get_initial_state(i2c_handle_t h, double* x0, double* sig_x0);
/* First action of the plan: cells[0].mu_u0_m / sig_u0_m (policy/mpc.py:166-167). */
int i2c_get_first_action(i2c_handle_t h, double* mu_u /*[B][du]*/, double* sig_u /*[B][du][du]*/);

/* Stand-alone sigma-point transform == QuadratureInference.forward / forward_gaussian
 * (inference/quadrature.py:27-58) for the registered env maps.  fn: 0 observe, 1 observe_terminal_x,
 * 2 forward (dynamics, adds nothing: S_noise is returned separately), 3 measure.
 * m[B][d], S[B][d][d] -> m_y[B][dy], S_y[B][dy][dy], S_xy[B][d][dy].  status[B] (may be NULL). */
int i2c_quadrature(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S,
                   double quad_alpha, double quad_beta, double quad_kappa, const double* env_par,
                   double* m_y, double* S_y, double* S_xy, int32_t* status, int32_t device);
/* Same transform with GaussHermiteQuadrature(degree) (exp_types.py:52-68; quadrature.py:132): degree^D tensor grid. */
int i2c_quadrature_gh(int32_t env, int32_t fn, int32_t n_problems, const double* m, const double* S, int32_t degree,
                      const double* env_par, double* m_y, double* S_y, double* S_xy, int32_t* status, int32_t device);
/* 1-D Gauss-Hermite rule used by the library: nodes[degree], weights[degree] = hermgauss weights / sqrt(pi). */
int i2c_gauss_hermite(int32_t degree, double* nodes, double* weights);

/* Batched stochastic closed-loop evaluation of the extracted controllers == BaseSim.run / batch_eval
 * (i2c/env.py:40-103; BaseKnownSim.forward :180-187) under TimeIndexedLinearGaussianPolicy /
 * ExpertTimeIndexedLinearGaussianPolicy (i2c/policy/linear.py:31-43, 73-90): n_rollouts roll-outs of every problem's
 * controller in one launch.  x_init[B][R][dx]; K[B][T][du][dx], k[B][T][du]; sigK[B][T][du][du] or NULL (deterministic
 * action); expert_mu[B][T][dx] + expert_lam[B][T][dx][dx] or NULL (plain linear policy), soft_expert: exp(-e) vs hard
 * gate; eta[B][R][T][dx] = process disturbances to add (parity runs) or NULL: drawn on the device (Philox, `seed`) from
 * N(0, sig_eta[dx][dx]); eps_u[B][R][T][du] standard normals for the action noise or NULL (device RNG).
 * Outputs: xu[B][R][T][dx+du] (state before the step and action), z[B][R][T][dz], z_term[B][R][dzt] (may be NULL),
 * x_final[B][R][dx] (state after the last step, may be NULL). */
int i2c_rollout(int32_t env, int32_t n_problems, int32_t n_rollouts, int32_t horizon, const double* x_init,
                const double* K, const double* k, const double* sigK, const double* expert_mu, const double* expert_lam,
                int32_t soft_expert, const double* eta, const double* eps_u, const double* sig_eta, uint64_t seed,
                const double* env_par, double* xu, double* z, double* z_term, double* x_final, int32_t device);

/* Snapshot / restore of the whole device state (deepcopy / dill pickling of the graph:
 * i2c.py:1392-1401, policy/mpc.py:24-26). */
int i2c_snapshot_bytes(i2c_handle_t h, size_t* bytes);
int i2c_snapshot(i2c_handle_t h, void* host_buf, size_t bytes);
int i2c_restore(i2c_handle_t h, const void* host_buf, size_t bytes);

/* Page-locked host buffers for the per-step traffic of the closed loop (measurements / applied actions in, actions out of
 * i2c_mpc_step; beliefs of i2c_set_initial_state_async): copies from / to them are truly asynchronous, whereas pageable NumPy
 * arrays (the reference passes those: policy/mpc.py:156-166) are staged by the driver, ~40 us per 128 KB array and step. */
int i2c_host_alloc(size_t bytes, void** out);
int i2c_host_free(void* p);

/* Introspection used by bench.py / tests. */
int i2c_kernel_launches(i2c_handle_t h, int64_t* n); /* kernels launched by this handle so far */
int i2c_last_run_ms(i2c_handle_t h, float* ms);      /* CUDA-event time of the last i2c_run kernel */
/* Measured fp64 FMA-pipe peak of the device in TFLOP/s (roofline denominator; 8 independent DFMA chains/thread). */
int i2c_dfma_peak(int32_t device, double* tflops);
/* The device math primitives (csrc/fastmath.cuh) on caller-provided arguments, for accuracy tests against libm:
 * fn 0 rsqrt, 1 reciprocal, 2 exp (x <= 0, throughput flavour), 3 exp (latency flavour), 4 sincos (branch-free),
 * 5 sincos (sequenced), 6 log-determinant accumulator log(x * x^2).  y1 receives the cosine for fn 4 / 5.  Replaces
 * nothing in the reference (NumPy libm calls: np.sin / np.cos env_def.py:273-291, np.linalg.cholesky quadrature.py:18,
 * scipy multivariate_normal.pdf i2c.py:369-374). */
int i2c_fastmath_probe(int32_t device, int32_t fn, int32_t n, const double* x, double* y0, double* y1);
const char* i2c_last_error(void);
const char* i2c_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* I2C_B200_H */
