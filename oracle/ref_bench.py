"""TEST / BENCH INFRASTRUCTURE ONLY -- timing harness around the UNMODIFIED reference (loaded through oracle/ref_shim.py
from /root/reference or its copy oracle/_ref): the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).

  em_worker    one process = a slice of the batched-pendulum problems of BASELINE configs[2], each run through the reference's
               own I2cGraph.learn_msgs (i2c/i2c.py:1238-1245), one problem after the other (the reference is single-problem)
  mpc_worker   one process = a few closed-loop roll-outs of BASELINE configs[4] through the reference's
               PartiallyObservedMpcPolicy.__call__ (i2c/policy/mpc.py:156-182) with the fp64 quadrotor restatement of
               oracle/envs.py substituted for the absent Box2D step (as tests/golden/make_golden.py: mpc does)
"""
import os
import time

import numpy as np


def _single_thread():
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"


def available():
    from oracle import ref_shim

    return ref_shim.available()


def em_worker(args):
    """args = (x0 [n,2], mu_u [n,T,1], hyper dict, steps, warmup) -> (seconds of the timed EM iterations, n, checksum)."""
    _single_thread()
    from oracle import ref_shim

    x0, mu_u, hyper, steps, warmup = args
    ns = ref_shim.load()
    graphs = []
    for b in range(x0.shape[0]):
        sys_ = ns.model.make_env_model("PendulumKnown", None)
        sys_.x0 = np.asarray(x0[b], float).reshape(-1, 1)
        g = ns.i2c.I2cGraph(sys_, mu_u.shape[1], hyper["Q"], hyper["R"], hyper["Q"], hyper["alpha"], hyper["tol"], mu_u[b],
                            hyper["sig_u"], None, None, ns.exp_types.CubatureQuadrature(1, 0, 0))
        for _ in range(warmup):
            g.learn_msgs()
        graphs.append(g)
    t0 = time.perf_counter()
    for g in graphs:
        for _ in range(steps):
            g.learn_msgs()
    dt = time.perf_counter() - t0
    return dt, len(graphs), float(sum(g.alpha for g in graphs))


def make_ref_quadrotor(ns):
    """The reference's model interface (BaseDef + BaseModelKnown, mpc_quad.py:219-386) around the fp64 quadrotor
    restatement (Box2D is absent: SURVEY.md 8c, DESIGN.md "parity unpinned")."""
    from oracle import envs as oenvs

    q = oenvs.Quadrotor()

    class QuadrotorDef(ns.env_def.BaseDef):
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

    class QuadrotorKnown(QuadrotorDef, ns.model.BaseModelKnown):
        pass

    return QuadrotorKnown, q


def mpc_setup(ns, feedforward=False, sig_zeta=None):
    from oracle import envs as oenvs

    QuadrotorKnown, q = make_ref_quadrotor(ns)
    W, H = oenvs.QUAD_W, oenvs.QUAD_H
    T, T_plan, mpc_iter = 100, 10, 2
    z_traj = np.zeros((T, 8))
    z_traj[:, 0] = np.linspace(W / 4, 3 * W / 4, T)
    z_traj[:, 1] = H / 2 + (H / 4) * np.sin(np.linspace(0, 2 * np.pi, T))
    z_traj[:, 2] = 2 * np.pi * np.heaviside(np.linspace(-1, 1, T), 1)
    Q, R = np.diag([1e3, 1e3, 1e3, 1, 1, 1]), np.diag([1e-3, 1e-3])
    model = QuadrotorKnown()
    model.sig_zeta = np.diag([1e-6] * 8) if sig_zeta is None else sig_zeta
    sig_u = 1e-2 * np.eye(2)
    g = ns.i2c.I2cGraph(sys=model, horizon=T_plan, Q=Q, R=R, Qf=Q / 1e3, alpha=1.0, alpha_update_tol=1.0,
                        mu_u=0.5 * q.gravity * np.ones((T_plan, 2)), sig_u=sig_u, mu_x_terminal=None, sig_x_terminal=None,
                        inference=ns.exp_types.CubatureQuadrature(1, 0, 0))
    g._propagate = True
    pol = ns.mpc.PartiallyObservedMpcPolicy(g, mpc_iter, sig_u, np.copy(z_traj))
    pol.set_control(feedforward=feedforward)
    pol.i2c.calibrate_alpha()
    pol.optimize(25, model.x0, model.sig_x0)
    pol.i2c.calibrate_alpha()
    return model, pol


def mpc_worker(args):
    """args = (n_rollouts, n_warm, n_timed, seed) -> (seconds spent in the timed policy calls, solves)."""
    _single_thread()
    from oracle import ref_shim

    n_roll, n_warm, n_timed, seed = args
    ns = ref_shim.load()
    rng = np.random.default_rng(seed)
    total, solves = 0.0, 0
    for _ in range(n_roll):
        model, pol = mpc_setup(ns)
        y0 = model.measure(model.x0[:, 0][None, :]).T
        u = np.zeros((2, 1))
        for t in range(n_warm + n_timed):
            y = y0 + 1e-3 * rng.normal(size=y0.shape)
            t0 = time.perf_counter()
            u = pol(t, y, u)
            dt = time.perf_counter() - t0
            u = model.clip_u(u.T).T
            if t >= n_warm:
                total += dt
                solves += 1
    return total, solves
