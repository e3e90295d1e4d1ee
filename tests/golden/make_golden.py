"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no tests / golden vectors, so parity is pinned
to outputs of the reference itself run here (Python 3.12, numpy 2.3.5, scipy 1.18.1; the
reference is imported unmodified through oracle/ref_shim.py).  Each .npz holds the inputs
needed to rebuild the problem plus per-cell message dumps / schedules.  The GPU box has no
/root/reference: tests only read the .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import envs as oenvs  # noqa: E402

ns = ref_shim.load()
I2cGraph = ns.i2c.I2cGraph
Cub = ns.exp_types.CubatureQuadrature
Lin = ns.exp_types.Linearize

FWD = ["mu_x0_f", "sig_x0_f", "mu_xu0_f", "sig_xu0_f", "mu_z0_f", "sig_z0_f", "mu_xu1_f", "sig_xu1_f",
       "mu_x3_f", "sig_x3_f", "J_dyn"]
BWD = ["mu_x3_m", "sig_x3_m", "mu_xu0_m", "sig_xu0_m", "mu_z0_m", "sig_z0_m", "K", "k", "sigK"]
PF = ["mu_x0_pf", "sig_x0_pf", "mu_xu0_pf", "sig_xu0_pf", "mu_z0_pf", "sig_z0_pf", "mu_x3_pf", "sig_x3_pf"]


def stack(g, attr):
    out = []
    for c in g.cells:
        a = np.asarray(getattr(c, attr), dtype=float)
        if a.ndim == 2 and a.shape[1] == 1 and attr.startswith(("mu_", "k", "nu_")):
            a = a[:, 0]
        elif a.ndim == 2 and a.shape[0] == 1 and attr.startswith("mu_z"):
            a = a[0]
        out.append(a)
    return np.stack(out)


def dump_iter(g, d, tag, names):
    for n in names:
        d[f"{tag}/{n}"] = stack(g, n)


def schedules(g, d):
    d["alphas"] = np.asarray(g.alphas, float)
    d["alphas_desired"] = np.asarray(g.alphas_desired, float)
    d["costs_m"] = np.asarray(g.costs_m, float)
    d["costs_m_var"] = np.asarray(g.costs_m_var, float)
    d["policy_entropy"] = np.asarray(g.policy_entropy, float)
    d["x_prior_entropy"] = np.asarray(g.x_prior_entropy, float)
    if g._propagate:
        d["alphas_pf"] = np.asarray(g.alphas_pf, float)
        d["costs_pf"] = np.asarray(g.costs_pf, float)
        d["costs_pf_var"] = np.asarray(g.costs_pf_var, float)
        d["cost_pf_min"] = np.asarray(g.cost_pf_min, float)
        d["propagate_entropy"] = np.asarray(g.propagate_entropy, float)
    if len(g.kl_terms):
        d["kl_terms"] = np.asarray(g.kl_terms, float)


def policy(g, d, tag="final"):
    K, k, s = g.get_local_linear_policy()
    d[f"{tag}/K"], d[f"{tag}/k"], d[f"{tag}/sigK"] = K, k, s


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(d)} arrays")


# ----------------------------------------------------------------------------- 1. quadrature KAT
def quad_kat():
    d = {}
    th = np.pi / 4
    T = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cov = T @ np.diag([0.5, 0.05]) @ T.T

    def func(x):  # inference/quadrature.py:97-104
        return np.concatenate((np.sin(1.5 * x[:, 1, None] + 1) + 0.1 * x[:, 0, None],
                               np.cos(1.5 * x[:, 1, None] + 1) + 0.1 * x[:, 0, None]), axis=1)

    q = ns.quadrature.QuadratureInference(Cub(1, 0, 0), 2)
    m, S = q.forward(func, np.zeros((2,)), cov)
    d["demo/cov"], d["demo/m"], d["demo/S"], d["demo/Sxy"] = cov, m[:, 0], S, q.sig_xy
    # every env's cost-feature map and dynamics through the reference quadrature object
    rng = np.random.default_rng(7)
    for key in ["LinearKnown", "LinearKnownMinimumEnergy", "PendulumKnown", "PendulumKnownActReg",
                "CartpoleKnown", "DoubleCartpoleKnown"]:
        sys_ = ns.model.make_env_model(key, None)
        n, dx = sys_.dim_xu, sys_.dim_x
        for rep in range(3):
            A = rng.normal(size=(n, n))
            S_in = 0.05 * (A @ A.T) / n + 1e-3 * np.eye(n)
            m_in = np.concatenate((sys_.x0[:, 0], [0.0] * sys_.dim_u)) + 0.3 * rng.normal(size=n)
            qo = ns.quadrature.QuadratureInference(Cub(1, 0, 0), n)
            mz, Sz = qo.forward(sys_.observe, m_in[:, None], S_in)
            qd = ns.quadrature.QuadratureInference(Cub(1, 0, 0), n)
            mx, Sx, Sn = qd.forward_gaussian(sys_.forward, m_in[:, None], S_in)
            t = f"{key}/{rep}"
            d[f"{t}/m_in"], d[f"{t}/S_in"] = m_in, S_in
            d[f"{t}/obs_m"], d[f"{t}/obs_S"], d[f"{t}/obs_Sxy"] = mz[:, 0], Sz, qo.sig_xy
            d[f"{t}/dyn_m"], d[f"{t}/dyn_S"], d[f"{t}/dyn_Sxy"], d[f"{t}/dyn_Sn"] = mx[:, 0], Sx, qd.sig_xy, Sn
            if key not in ("PendulumKnownActReg",):
                qt = ns.quadrature.QuadratureInference(Cub(1, 0, 0), dx)
                mt, St = qt.forward(sys_.observe_terminal_x, m_in[:dx, None], S_in[:dx, :dx])
                d[f"{t}/term_m"], d[f"{t}/term_S"], d[f"{t}/term_Sxy"] = mt[:, 0], St, qt.sig_xy
    # general (alpha, beta, kappa) weights
    for (a, b, k) in [(1, 0, 0), (0.5, 2.0, 1.0), (1.0, 2.0, 0.5)]:
        for dim in (2, 3, 5, 7, 8):
            sf, wm, ws = Cub(a, b, k).weights(dim)
            d[f"weights/{a}_{b}_{k}/{dim}"] = np.concatenate(([sf], wm, ws))
    save("quadrature_kat", d)


# ----------------------------------------------------------------------------- 2. EM runs
def em_run(name, env_key, T, Q, R, Qf, alpha, tol, mu_u, sig_u, n_dump, n_total, mu_x_term=None, sig_x_term=None,
           propagate=False, expert=True, x0=None, inference=None):
    sys_ = ns.model.make_env_model(env_key, None)
    if x0 is not None:
        sys_.x0 = np.asarray(x0, float).reshape(-1, 1)
    g = I2cGraph(sys_, T, Q, R, Qf, alpha, tol, mu_u, sig_u, mu_x_term, sig_x_term, inference or Cub(1, 0, 0))
    d = dict(env=env_key, T=T, Q=np.zeros(0) if Q is None else Q, R=R, Qf=np.zeros(0) if Qf is None else Qf,
             alpha0=alpha, tol=tol, mu_u=mu_u, sig_u=sig_u, x0=sys_.x0[:, 0],
             mu_x_term=np.zeros(0) if mu_x_term is None else np.asarray(mu_x_term, float).reshape(-1),
             sig_x_term=np.zeros(0) if sig_x_term is None else sig_x_term,
             propagate=propagate, expert=expert, n_dump=n_dump, n_total=n_total)
    if propagate:
        g._propagate = True
        for c in g.cells:
            c.use_expert_controller = expert
        g.propagate()
        dump_iter(g, d, "it0", PF)
    for it in range(1, n_total + 1):
        g.learn_msgs()
        if it <= n_dump:
            dump_iter(g, d, f"it{it}", FWD + BWD + (PF if propagate else []))
            if g.cells[-1].mu_z3_m is not None:
                d[f"it{it}/mu_z3_m"] = np.asarray(g.cells[-1].mu_z3_m)[:, 0]
                d[f"it{it}/sig_z3_m"] = np.asarray(g.cells[-1].sig_z3_m)
    schedules(g, d)
    policy(g, d)
    d["final/mu_xu0_m"] = stack(g, "mu_xu0_m")
    d["final/sig_xu0_m"] = stack(g, "sig_xu0_m")
    save(name, d)


def em_runs():
    # config 1: scripts/experiments/pendulum_known_quad.py, seed 0, T=100, 200 iterations
    exp = ref_shim.load_experiment("pendulum_known_quad", 0)
    I = exp.INFERENCE
    em_run("pendulum_known_quad_seed0", exp.ENVIRONMENT, exp.N_DURATION, I.Q, I.R, I.Qf, I.alpha,
           I.alpha_update_tol, I.mu_u, I.sig_u, n_dump=3, n_total=200)
    rng = np.random.default_rng(11)
    # perturbed initial state pendulum, T=200 (config-3 shaped single problem)
    em_run("pendulum_T200_x0pert", "PendulumKnown", 200, I.Q, I.R, I.Qf, 100.0, 0.0,
           1e-2 * rng.normal(size=(200, 1)), I.sig_u, n_dump=2, n_total=30,
           x0=np.array([np.pi, 0.0]) + np.array([0.3, 0.5]) * rng.normal(size=2))
    # cart-pole (hyper-parameters of cartpole_known_quad.py:23-34), shorter horizon
    em_run("cartpole_T120", "CartpoleKnown", 120, np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]),
           np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), 80.0, 0.0, 1e-2 * rng.normal(size=(120, 1)), 1.0 * np.eye(1),
           n_dump=2, n_total=20)
    # double cart-pole (double_cartpole_known_cq.py:23-39)
    sf = 1e-3
    Qd = sf * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0])
    em_run("double_cartpole_T80", "DoubleCartpoleKnown", 80, Qd, sf * np.diag([0.1]), Qd, 0.05, 0.99,
           1e-2 * rng.normal(size=(80, 1)), np.eye(1), n_dump=2, n_total=12)
    # covariance control + propagate (pendulum_known_act_reg_quad.py:22-33, nonlinear_covariance_control.py:81-115)
    em_run("pendulum_actreg_covctrl_T100", "PendulumKnownActReg", 100, None, np.diag([1.0]), None, 300.0, 1.0,
           np.zeros((100, 1)), 0.5 * np.eye(1), n_dump=3, n_total=15, mu_x_term=np.array([0.0, 0.0]),
           sig_x_term=np.diag([1e-3, 1e-3]), propagate=True, expert=False)
    # double cart-pole covariance control with the expert down-weighting on (config-4 analogue, short)
    em_run("double_cartpole_covctrl_T50", "DoubleCartpoleKnown", 50, Qd, sf * np.diag([0.1]), Qd, 0.05, 0.99,
           1e-2 * rng.normal(size=(50, 1)), np.eye(1), n_dump=2, n_total=8, mu_x_term=np.zeros(6),
           sig_x_term=np.diag([0.01, 0.005, 0.005, 0.05, 0.05, 0.05]), propagate=True, expert=False)
    em_run("pendulum_propagate_expert_T50", "PendulumKnown", 50, I.Q, I.R, I.Qf, 100.0, 0.0,
           1e-2 * rng.normal(size=(50, 1)), I.sig_u, n_dump=3, n_total=6, propagate=True, expert=True)
    # linear minimum-energy system with cubature (well conditioned linear case)
    em_run("linear_minenergy_cubature_T40", "LinearKnownMinimumEnergy", 40, None, np.diag([1.0]), None, 10.0, 0.5,
           1e-2 * rng.normal(size=(40, 1)), 1e1 * np.eye(1), n_dump=2, n_total=6)


def gauss_hermite():
    """GaussHermiteQuadrature (exp_types.py:52-68) through the unmodified reference."""
    GH = ns.exp_types.GaussHermiteQuadrature
    d = {}
    # the reference's own demo (inference/quadrature.py:86-104, 132): 2-D, degree 4
    th = np.pi / 4
    T = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cov = T @ np.diag([0.5, 0.05]) @ T.T
    rng = np.random.default_rng(17)
    for key, deg in [("PendulumKnown", 4), ("PendulumKnown", 3), ("LinearKnown", 5), ("CartpoleKnown", 3)]:
        sys_ = ns.model.make_env_model(key, None)
        n, dx = sys_.dim_xu, sys_.dim_x
        A = rng.normal(size=(n, n))
        S_in = 0.05 * (A @ A.T) / n + 1e-3 * np.eye(n)
        m_in = np.concatenate((sys_.x0[:, 0], [0.0] * sys_.dim_u)) + 0.3 * rng.normal(size=n)
        qo = ns.quadrature.QuadratureInference(GH(deg), n)
        mz, Sz = qo.forward(sys_.observe, m_in[:, None], S_in)
        qd = ns.quadrature.QuadratureInference(GH(deg), n)
        mx, Sx, Sn = qd.forward_gaussian(sys_.forward, m_in[:, None], S_in)
        qt = ns.quadrature.QuadratureInference(GH(deg), dx)
        mt, St = qt.forward(sys_.observe_terminal_x, m_in[:dx, None], S_in[:dx, :dx])
        t = f"{key}/{deg}"
        d[f"{t}/m_in"], d[f"{t}/S_in"] = m_in, S_in
        d[f"{t}/obs_m"], d[f"{t}/obs_S"], d[f"{t}/obs_Sxy"] = mz[:, 0], Sz, qo.sig_xy
        d[f"{t}/dyn_m"], d[f"{t}/dyn_S"], d[f"{t}/dyn_Sxy"], d[f"{t}/dyn_Sn"] = mx[:, 0], Sx, qd.sig_xy, Sn
        d[f"{t}/term_m"], d[f"{t}/term_S"], d[f"{t}/term_Sxy"] = mt[:, 0], St, qt.sig_xy
    for deg in (1, 2, 3, 4, 7):
        sf, wm, ws = GH(deg).weights(2)
        d[f"weights/{deg}"] = np.concatenate(([sf], wm))
        d[f"pts/{deg}"] = GH(deg).pts(2)
    save("gauss_hermite_kat", d)
    exp = ref_shim.load_experiment("pendulum_known_quad", 0)
    I = exp.INFERENCE
    em_run("pendulum_gh3_T40", "PendulumKnown", 40, I.Q, I.R, I.Qf, 100.0, 0.0, 1e-2 * rng.normal(size=(40, 1)), I.sig_u,
           n_dump=2, n_total=8, inference=GH(3))
    em_run("pendulum_gh4_propagate_T20", "PendulumKnown", 20, I.Q, I.R, I.Qf, 100.0, 0.5, 1e-2 * rng.normal(size=(20, 1)),
           I.sig_u, n_dump=2, n_total=4, inference=GH(4), propagate=True, expert=False)


# ----------------------------------------------------------------------------- 3. LQR / Linearize
def lqr():
    """scripts/lqr_compare.py:120-176 logic (no plots)."""
    exp = ref_shim.load_experiment("linear_known", 0)
    model = ns.model.make_env_model(exp.ENVIRONMENT, exp.MODEL)
    model.xag = 10 * np.ones((2, 1))
    model.zg_term = 10 * np.ones((2, 1))
    model.a = model.xag - model.A @ model.xag
    I = exp.INFERENCE
    H = exp.N_DURATION
    x_lqr, u_lqr, K_lqr, k_lqr, cost_lqr, P, p = ns.utils.finite_horizon_lqr(
        H, model.A, model.a[:, 0], model.B, I.Q, I.R, model.x0[:, 0], model.xag[:, 0], np.zeros((1,)), 2, 1)
    g = I2cGraph(sys=model, horizon=H, Q=I.Q, R=I.R, Qf=I.Qf, alpha=1e-5, alpha_update_tol=I.alpha_update_tol,
                 mu_u=np.zeros((H, 1)), sig_u=1e2 * np.eye(1), mu_x_terminal=None, sig_x_terminal=None,
                 inference=I.inference, res_dir=None)
    for c in g.cells:
        c.state_action_independence = True
    g._forward_backward_msgs()
    d = dict(H=H, Q=I.Q, R=I.R, Qf=I.Qf, A=model.A, B=model.B, a=model.a[:, 0], xag=model.xag[:, 0],
             x0=model.x0[:, 0], K_lqr=K_lqr, k_lqr=k_lqr, x_lqr=x_lqr, u_lqr=u_lqr, P=P, p=p, cost_lqr=cost_lqr)
    dump_iter(g, d, "fb", ["mu_xu0_f", "sig_xu0_f", "mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn",
                           "mu_x3_m", "sig_x3_m", "mu_xu0_m", "sig_xu0_m", "mu_z0_m", "sig_z0_m", "K", "k", "sigK"])
    g._backward_ricatti_msgs()
    dump_iter(g, d, "ric", ["K", "k", "lambda_x3_b", "nu_x3_b", "lambda_x0_b", "nu_x0_b"])
    save("lqr_linearize", d)

    # linear covariance control (linear_known_covariance_control.py + linear_gaussian_covariance_control.py:91-125)
    exp = ref_shim.load_experiment("linear_known_covariance_control", 0)
    model = ns.model.make_env_model(exp.ENVIRONMENT, exp.MODEL)
    I = exp.INFERENCE
    g = I2cGraph(sys=model, horizon=exp.N_DURATION, Q=I.Q, R=I.R, Qf=I.Qf, alpha=I.alpha,
                 alpha_update_tol=I.alpha_update_tol, mu_u=I.mu_u, sig_u=I.sig_u, mu_x_terminal=I.mu_x_term,
                 sig_x_terminal=I.sig_x_term, inference=I.inference, res_dir=None)
    for c in g.cells:
        c.use_expert_controller = False
    g._propagate = True
    d = dict(T=exp.N_DURATION, R=I.R, alpha0=I.alpha, tol=I.alpha_update_tol, mu_u=I.mu_u, sig_u=I.sig_u,
             mu_x_term=np.asarray(I.mu_x_term, float).reshape(-1), sig_x_term=I.sig_x_term)
    for it in range(1, 6):
        g.learn_msgs()
        if it <= 2:
            dump_iter(g, d, f"it{it}", ["mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "mu_xu0_m", "sig_xu0_m",
                                        "K", "k", "sigK", "mu_x3_pf", "sig_x3_pf"])
    schedules(g, d)
    policy(g, d)
    save("linear_covctrl_linearize", d)


def linearize_nonlinear():
    """Linearize inference on the nonlinear envs (experiments pendulum_known.py, cartpole_known.py,
    double_cartpole_known(_lin).py): the reference's I2cCell._forward_msgs_linearize / _backward_msgs_linearize with the
    Jacobians of its OWN env_autograd.py dynamics.  autograd is absent from this image; oracle/ref_shim.py supplies
    autograd.jacobian as the complex-step derivative (exact to rounding for these analytic functions).  Same inputs as
    tests/test_gpu_lqr.py::test_linearize_nonlinear_envs, problem 0."""
    cases = [
        ("pendulum_linearize_T40", "PendulumKnown", np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, [0.3, 0.5], 40),
        ("cartpole_linearize_T40", "CartpoleKnown", np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]), 80.0, 0.05, 40),
        ("double_cartpole_linearize_T30", "DoubleCartpoleKnown",
         1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.02, 30),
    ]
    for name, env, Q, R, alpha, xs, T in cases:
        rng = np.random.default_rng(4)
        e = oenvs.make(env)
        B = 16  # the GPU test draws a batch of 16 with this seed; the golden is its problem 0
        x0 = e.x0 + np.asarray(xs) * rng.normal(size=(B, e.dim_x))
        mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
        em_run(name, env, T, Q, R, Q, alpha, 0.5, mu_u[0], np.eye(e.dim_u), n_dump=3, n_total=3, x0=x0[0], inference=Lin())


# ----------------------------------------------------------------------------- 4. MPC (quadrotor, fp64 stand-in)
def mpc():
    """policy/mpc.py:115-182 driven as mpc_quad.py:538-652 does, with the fp64 quadrotor restatement
    (oracle/envs.py:Quadrotor) substituted for the Box2D step (absent here) on the reference side."""
    BaseDef = ns.env_def.BaseDef
    BaseModelKnown = ns.model.BaseModelKnown
    q = oenvs.Quadrotor()

    class QuadrotorDef(BaseDef):
        name = "2D Quadrator"
        dim_x, dim_u, dim_z, dim_y = 6, 2, 8, 8
        dim_z_term = 6
        x0 = q.x0[:, None].copy()
        sig_x0 = q.sig_x0.copy()
        sig_eta = q.sig_eta.copy()
        sig_zeta = None
        xag = q.zg_term[:, None].copy()
        zg_term = xag
        xu_lim = np.array([[-np.inf] * 6 + [0.0, 0.0], [np.inf] * 6 + [30.0, 30.0]])

        def dynamics(self, xu):
            return oenvs.Quadrotor.dynamics(xu)

        @staticmethod
        def observe(xu):
            return xu

        @staticmethod
        def observe_terminal(x):
            return x

        @staticmethod
        def measure(x):
            return oenvs.Quadrotor.measure(x)

    class QuadrotorKnown(QuadrotorDef, BaseModelKnown):
        pass

    W, H = oenvs.QUAD_W, oenvs.QUAD_H
    T, T_plan, mpc_iter = 100, 10, 2
    z_traj = np.zeros((T, 8))
    z_traj[:, 0] = np.linspace(W / 4, 3 * W / 4, T)
    z_traj[:, 1] = H / 2 + (H / 4) * np.sin(np.linspace(0, 2 * np.pi, T))
    z_traj[:, 2] = 2 * np.pi * np.heaviside(np.linspace(-1, 1, T), 1)
    Q = np.diag([1e3, 1e3, 1e3, 1, 1, 1])
    R = np.diag([1e-3, 1e-3])
    Qf = Q / 1e3
    n_steps = 14
    for feedforward in (True, False):
        for low_noise in (True, False):
            model = QuadrotorKnown()
            sig_zeta = np.diag([1e-6] * 8) if low_noise else np.diag([1e-6] * 2 + [5e-5] * 2 + [1] * 4)
            model.sig_zeta = sig_zeta
            u_init = 0.5 * q.gravity * np.ones((T_plan, 2))
            sig_u = 1e-2 * np.eye(2)
            g = I2cGraph(sys=model, horizon=T_plan, Q=Q, R=R, Qf=Qf, alpha=1.0, alpha_update_tol=1.0, mu_u=u_init,
                         sig_u=sig_u, mu_x_terminal=None, sig_x_terminal=None, inference=Cub(1, 0, 0))
            g._propagate = True
            pol = ns.mpc.PartiallyObservedMpcPolicy(g, mpc_iter, sig_u, np.copy(z_traj))
            pol.set_control(feedforward=feedforward)
            rng = np.random.default_rng(5 + 2 * int(feedforward) + int(low_noise))
            eta = rng.multivariate_normal(np.zeros(6), model.sig_eta, n_steps)
            zeta = rng.multivariate_normal(np.zeros(8), sig_zeta, n_steps + 1)
            d = dict(z_traj=z_traj, Q=Q, R=R, Qf=Qf, sig_zeta=sig_zeta, u_init=u_init, sig_u=sig_u, eta=eta, zeta=zeta,
                     T_plan=T_plan, mpc_iter=mpc_iter, feedforward=feedforward)
            pol.i2c.calibrate_alpha()
            d["alpha_cal1"] = pol.i2c.alpha
            pol.optimize(25, model.x0, model.sig_x0)
            pol.i2c.calibrate_alpha()
            d["alpha_cal2"] = pol.i2c.alpha
            d["warm/mu_xu0_m"] = stack(pol.i2c, "mu_xu0_m")
            d["warm/sig_xu0_m"] = stack(pol.i2c, "sig_xu0_m")
            d["warm/K"] = stack(pol.i2c, "K")
            x = model.x0[:, 0].copy()
            y = model.measure(x[None, :]).T + zeta[0][:, None]
            u = np.zeros((2, 1))
            us, mus, covs, xs, ys, plan = [], [], [], [], [], []
            for t in range(n_steps):
                ys.append(y[:, 0].copy())
                u = pol(t, y, u)
                mus.append(pol.mus[-1][:, 0].copy())
                covs.append(pol.covars[-1].copy())
                plan.append(pol.xu_history[-1][:, :, 0].copy())
                u = model.clip_u(u.T).T
                us.append(u[:, 0].copy())
                xs.append(x.copy())
                x = model.dynamics(np.concatenate((x, u[:, 0]))[None, :])[0] + eta[t]
                y = model.measure(x[None, :]).T + zeta[t + 1][:, None]
            d["u"], d["mu"], d["covar"], d["x"], d["y"], d["plan"] = map(np.asarray, (us, mus, covs, xs, ys, plan))
            save(f"mpc_quadrotor_{'ff' if feedforward else 'fb'}_{'low' if low_noise else 'high'}", d)


def evaluators():
    """TrajectoryEvaluator / StochasticTrajectoryEvaluator (i2c/utils.py:103-265): the consumers of the roll-outs in
    scripts/i2c_run.py:66-106, and the cost_*.npy files they write."""
    import tempfile

    rng = np.random.default_rng(5)
    T, R, dz, dzt, dx = 30, 12, 4, 3, 2
    A = rng.normal(size=(dz, dz))
    W = A @ A.T
    Af = rng.normal(size=(dx, dx))
    Wf = Af @ Af.T
    sg, sg_term = rng.normal(size=dz), rng.normal(size=dzt)
    d = dict(W=W, Wf=Wf, sg=sg, sg_term=sg_term, dim_x=dx)
    ev = ns.utils.StochasticTrajectoryEvaluator(W, Wf, sg, sg_term, dx)
    det = ns.utils.TrajectoryEvaluator(W, Wf, sg, sg_term, dx)
    for k in range(3):
        trajs, terms = rng.normal(size=(R, T, dz)), rng.normal(size=(R, dzt))
        plan, plan_term = rng.normal(size=(T, dz)), rng.normal(size=(1, dzt))
        ev.eval(trajs, terms, plan, plan_term)
        det.eval(trajs[0], terms[:1], plan, plan_term)
        d[f"{k}/trajs"], d[f"{k}/terms"], d[f"{k}/plan"], d[f"{k}/plan_term"] = trajs, terms, plan, plan_term
    for name in ["mu_actual_cost", "max_actual_cost", "min_actual_cost", "actual_cost_10", "actual_cost_90", "planned_cost"]:
        d[f"stoch/{name}"] = np.asarray(getattr(ev, name), float)
    d["det/actual_cost"], d["det/planned_cost"] = np.asarray(det.actual_cost, float), np.asarray(det.planned_cost, float)
    with tempfile.TemporaryDirectory() as tmp:
        ev.save("x", tmp)
        det.save("y", tmp)
        for f in sorted(os.listdir(tmp)):
            d[f"file/{f}"] = np.load(os.path.join(tmp, f))
    save("evaluator_kat", d)


def likelihood():
    """Likelihood diagnostics (I2cGraph.calc_likelihood, i2c.py:1135-1164): never called by the reference's scripts,
    pinned here by calling it after every EM iteration."""
    exp = ref_shim.load_experiment("pendulum_known_quad", 0)
    I = exp.INFERENCE
    rng = np.random.default_rng(3)
    T = 30
    mu_u = 1e-2 * rng.normal(size=(T, 1))
    sys_ = ns.model.make_env_model("PendulumKnown", None)
    g = I2cGraph(sys_, T, I.Q, I.R, I.Qf, 100.0, 0.5, mu_u, I.sig_u, None, None, Cub(1, 0, 0))
    for _ in range(4):
        g.learn_msgs()
        g.calc_likelihood()
    d = dict(T=T, Q=I.Q, R=I.R, Qf=I.Qf, alpha0=100.0, tol=0.5, mu_u=mu_u, sig_u=I.sig_u,
             likelihoods=np.asarray(g.likelihoods, float), likelihoods_xu=np.asarray(g.likelihoods_xu, float),
             likelihoods_z=np.asarray(g.likelihoods_z, float), risk=np.asarray(g.risk, float).reshape(-1))
    vals = [3.0, 2.5, 2.0, 2.2, 1.9, 1.8, 1.7]
    d["minima"] = np.array([[float(x) if x is not None else -1.0 for x in
                             (I2cGraph.list_minima(vals[:n], 2, 2), I2cGraph.list_minima(vals[:n], 2, 3))] for n in range(1, 8)])
    save("likelihood_kat", d)


def rollouts():
    """Pin for oracle/rollout.py: the reference's BaseSim.run (i2c/env.py:40-74) with BaseKnownSim.forward (:180-187)
    under TimeIndexedLinearGaussianPolicy / ExpertTimeIndexedLinearGaussianPolicy (policy/linear.py:31-90), global NumPy RNG
    seeded.  Every disturbance the reference draws (numpy.random.multivariate_normal in i2c.env and i2c.policy.linear) is
    logged, so that the oracle / the CUDA kernel can be fed the very same realisations."""
    import importlib

    env_mod = importlib.import_module("i2c.env")
    lin_mod = importlib.import_module("i2c.policy.linear")
    log = []
    real_mvn = np.random.multivariate_normal

    def logged(mean, cov, size=None):
        out = real_mvn(mean, cov, size)
        log.append(np.asarray(out, float).reshape(-1) - np.asarray(mean, float).reshape(-1))
        return out

    env_mod.mvn = logged
    lin_mod.mvn = logged
    d = {}
    for env_key, exp_name, T in [("PendulumKnown", "pendulum_known_quad", 40), ("CartpoleKnown", "cartpole_known_quad", 30)]:
        exp = ref_shim.load_experiment(exp_name, 0)
        I = exp.INFERENCE
        exp.N_DURATION = T
        sim = env_mod.make_env(exp)
        sys_ = ns.model.make_env_model(env_key, None)
        rng = np.random.default_rng(9)
        mu_u = 1e-2 * rng.normal(size=(T, sys_.dim_u))
        g = I2cGraph(sys_, T, I.Q, I.R, I.Qf, I.alpha, 0.0, mu_u, I.sig_u, None, None, Cub(1, 0, 0))
        for _ in range(6):
            g.learn_msgs()
        K, k, sk = g.get_local_linear_policy()
        pol = lin_mod.TimeIndexedLinearGaussianPolicy(I.sig_u, T, sys_.dim_u, sys_.dim_x)
        pol.write(K, k, sk)
        Ke, ke, ske, mue, lame = g.get_local_expert_linear_policy()
        d[f"{env_key}/T"], d[f"{env_key}/K"], d[f"{env_key}/k"], d[f"{env_key}/sigK"] = T, K, k, sk
        d[f"{env_key}/x0"] = np.asarray(sim.x0, float).reshape(-1)
        d[f"{env_key}/ex_K"], d[f"{env_key}/ex_k"], d[f"{env_key}/ex_mu"], d[f"{env_key}/ex_lam"] = Ke, ke, mue, lame
        runs = [("det_env_det_pol", True, True, pol), ("noisy_env_det_pol", False, True, pol),
                ("noisy_env_noisy_pol", False, False, pol)]
        for soft in (True, False):
            ep = lin_mod.ExpertTimeIndexedLinearGaussianPolicy(I.sig_u, T, sys_.dim_u, sys_.dim_x, soft=soft)
            ep.write(Ke, ke, ske, mue, lame)
            runs.append((f"expert_{'soft' if soft else 'hard'}", False, True, ep))
        for tag, env_det, pol_det, policy in runs:
            np.random.seed(123)
            sim.deterministic = env_det
            log.clear()
            xt, yt, zt, z_term = sim.run(policy, deterministic=pol_det)
            draws = [np.array(v) for v in log]
            # order of the draws inside one step: policy first (if stochastic), then the environment (env.py:57-61)
            eta = np.zeros((T, sys_.dim_x))
            eu = np.zeros((T, sys_.dim_u))
            it = iter(draws)
            for t in range(T):
                if not pol_det:
                    eu[t] = next(it)
                if not env_det:
                    eta[t] = next(it)
            assert next(it, None) is None
            d[f"{env_key}/{tag}/xt"], d[f"{env_key}/{tag}/zt"], d[f"{env_key}/{tag}/z_term"] = xt, zt, z_term
            d[f"{env_key}/{tag}/eta"], d[f"{env_key}/{tag}/u_noise"] = eta, eu
    env_mod.mvn = real_mvn
    lin_mod.mvn = real_mvn
    save("rollout_kat", d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["quad", "em", "lqr", "linnl", "mpc", "gh", "eval", "ll", "roll"]
    if "quad" in which:
        quad_kat()
    if "em" in which:
        em_runs()
    if "lqr" in which:
        lqr()
    if "linnl" in which:
        linearize_nonlinear()
    if "mpc" in which:
        mpc()
    if "gh" in which:
        gauss_hermite()
    if "eval" in which:
        evaluators()
    if "ll" in which:
        likelihood()
    if "roll" in which:
        rollouts()
