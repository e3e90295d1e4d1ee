"""Seeded problem factories shared by the GPU tests (hyper-parameters of the reference's experiment files:
experiments/pendulum_known_quad.py:22-33, cartpole_known_quad.py, double_cartpole_known_cq.py:23-39, mpc_quad.py:558-594)."""
import numpy as np

HYP = {
    "PendulumKnown": dict(Q=np.diag([1.0, 100.0, 1.0]), R=np.diag([2.0]), alpha=100.0, tol=0.0, sig_u=2.0, xs=[0.3, 0.5]),
    "CartpoleKnown": dict(Q=np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), R=np.diag([1.0]), alpha=80.0, tol=0.0, sig_u=1.0, xs=0.05),
    "DoubleCartpoleKnown": dict(Q=1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), R=1e-4 * np.eye(1),
                                alpha=0.05, tol=0.99, sig_u=1.0, xs=0.02),
    "Quadrotor": dict(Q=np.diag([1e3, 1e3, 1e3, 1, 1, 1]), R=np.diag([1e-3, 1e-3]), alpha=1.0, tol=1.0, sig_u=1e-2, xs=0.01),
}


def make_case(m, env, B, T, seed=11, **kw):
    h = HYP[env]
    e = m.envs.make(env)
    rng = np.random.default_rng(seed)
    x0 = e.x0 + np.asarray(h["xs"]) * rng.normal(size=(B, e.dim_x))
    mu0 = 0.5 * 9.81 * m.envs.QUAD_MASS if env == "Quadrotor" else 0.0
    mu_u = mu0 + 1e-2 * rng.normal(size=(B, T, e.dim_u))
    Qf = h["Q"] / (1e3 if env == "Quadrotor" else 1.0)
    return m.BatchedI2c(env, B, T, h["Q"], h["R"], Qf, h["alpha"], h["tol"], mu_u, h["sig_u"] * np.eye(e.dim_u), x0=x0,
                        max_iters=8, **kw)
