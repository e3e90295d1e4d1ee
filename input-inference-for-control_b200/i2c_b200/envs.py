"""Host-side constants of the registered environments (what I2cGraph reads off ``sys``: x0, sig_x0, sig_eta,
zg, zg_term and the dimensions).  Values restate the reference's definitions (paths relative to the
reference root); the dynamics / cost-feature maps themselves live in csrc/envs.cuh."""
import numpy as np

QUAD_W, QUAD_H = 600 / 30.0, 400 / 30.0
QUAD_DX, QUAD_DY = QUAD_W / 25, QUAD_H / 100
QUAD_MASS = 5.0 * (2 * QUAD_DX) * (2 * QUAD_DY)


class EnvConst:
    def __init__(self, name, dx, du, dz, dzt, x0, sig_x0, sig_eta, zg, zg_term, has_term=True, dy=0, u_lim=None):
        self.name, self.dim_x, self.dim_u, self.dim_z, self.dim_z_term, self.dim_y = name, dx, du, dz, dzt, dy
        self.x0 = np.asarray(x0, float)
        self.sig_x0 = np.asarray(sig_x0, float)
        self.sig_eta = np.asarray(sig_eta, float)
        self.zg = np.asarray(zg, float)
        self.zg_term = np.asarray(zg_term, float)
        self.has_term = has_term
        self.u_lim = u_lim

    @property
    def dim_xu(self):
        return self.dim_x + self.dim_u


def linear_params(A, B, a):
    """per-problem parameter vector of the linear envs: [A row-major (4), B (2), a (2)]"""
    A, B, a = np.asarray(A, float), np.asarray(B, float), np.asarray(a, float)
    lead = A.shape[:-2]
    Bb = np.broadcast_to(B.reshape(B.shape[:-2] + (2,)) if B.shape[-1] == 1 else B, lead + (2,))
    return np.concatenate((A.reshape(lead + (4,)), Bb, np.broadcast_to(a, lead + (2,))), axis=-1)


def _linear():
    # env_def.py:139-191
    A = np.array([[1.1, 0.0], [0.1, 1.1]])
    xg = np.array([1.0, -1.0])
    c = EnvConst("LinearKnown", 2, 1, 3, 2, [5.0, 5.0], 1e-20 * np.eye(2), 1e-20 * np.eye(2), [1.0, -1.0, 0.0], xg)
    c.A, c.B, c.a = A, np.array([[0.1], [0.0]]), xg - A @ xg
    return c


def _linear_min_energy():
    # env_def.py:194-230
    A = np.array([[1.05, 0.0], [0.05, 1.01]])
    g = np.array([-5.0, -5.0])
    c = EnvConst("LinearKnownMinimumEnergy", 2, 1, 1, 2, [5.0, 5.0], np.diag([1e-1, 5e0]), np.diag([1e-1, 1e-2]), [0.0], g)
    c.A, c.B, c.a = A, np.array([[0.1], [0.0]]), g - A @ g
    return c


REGISTRY = {
    "LinearKnown": _linear,
    "LinearKnownMinimumEnergy": _linear_min_energy,
    # env_def.py:233-309
    "PendulumKnown": lambda: EnvConst("PendulumKnown", 2, 1, 4, 3, [np.pi, 0.0], 1e-5 * np.eye(2), np.diag([1e-5, 1e-5]),
                                      [0.0, 1.0, 0.0, 0.0], [0.0, 1.0, 0.0], u_lim=(-2.0, 2.0)),
    # env_def.py:312-346
    "PendulumKnownActReg": lambda: EnvConst("PendulumKnownActReg", 2, 1, 1, 1, [np.pi, 0.0], 1e-5 * np.eye(2),
                                            np.diag([1e-5, 1e-5]), [0.0], [0.0], has_term=False, u_lim=(-2.0, 2.0)),
    # env_def.py:491-612
    "CartpoleKnown": lambda: EnvConst("CartpoleKnown", 4, 1, 6, 5, [0.0, np.pi, 0.0, 0.0], 1e-5 * np.eye(4),
                                      np.diag([1e-8] * 4), [0.0, 0.0, 1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0, 0.0],
                                      u_lim=(-5.0, 5.0)),
    # env_def.py:615-761
    "DoubleCartpoleKnown": lambda: EnvConst("DoubleCartpoleKnown", 6, 1, 9, 8, [0.0, np.pi, np.pi, 0.0, 0.0, 0.0],
                                            1e-6 * np.eye(6), np.diag([1e-6] * 6),
                                            [0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0],
                                            [0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0], u_lim=(-10.0, 10.0)),
    # scripts/mpc_state_est/mpc_quad.py:219-260
    "Quadrotor": lambda: EnvConst("Quadrotor", 6, 2, 8, 6, [QUAD_W / 4, QUAD_H / 2, 0, 0, 0, 0], 1e-5 * np.eye(6),
                                  np.diag([1e-6] * 3 + [1e-4] * 3), [3 * QUAD_W / 4, QUAD_H / 2, 0, 0, 0, 0, 0, 0],
                                  [3 * QUAD_W / 4, QUAD_H / 2, 0, 0, 0, 0], dy=8, u_lim=(0.0, 30.0)),
}


def make(name):
    if name not in REGISTRY:
        raise KeyError(f"environment {name!r} is not registered with the CUDA path (no CPU fallback); "
                       f"known: {sorted(REGISTRY)}")
    return REGISTRY[name]()
