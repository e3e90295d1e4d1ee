"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's environment
definitions (dynamics, cost-feature maps, constants) for the CPU oracle.

Every function works on arrays with arbitrary leading dimensions ``(..., d)`` so the
oracle can be batched over problems.  Citations are into /root/reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product never does.
"""
import numpy as np


class EnvSpec:
    """Plain container mirroring the attributes i2c reads from ``sys``
    (env_def.py BaseDef + model.py BaseModelKnown)."""

    name = None
    dim_x = dim_u = dim_z = dim_z_term = None
    has_terminal_obs = True

    @property
    def dim_xu(self):
        return self.dim_x + self.dim_u

    def forward(self, xu):
        """model.py:154-156 -- mean step + constant process noise (broadcast by caller)."""
        return self.dynamics(xu), self.sig_eta

    def observe_terminal_x(self, x):  # model.py:100-101
        return self.observe_terminal(x)

    # ---- linearisations for the Linearize inference on nonlinear envs.  The reference takes the dynamics Jacobian
    # with autograd (env_autograd.py:22,57,170; model.py:158-164) -- absent from this image -- here: the complex-step
    # derivative of the restated dynamics (no truncation, no cancellation: the exact derivative to rounding, like autograd's).
    # Pinned by tests/golden/*_linearize_*.npz: the unmodified reference run with the same complex-step stand-in for
    # autograd.jacobian (oracle/ref_shim.py).  The cost-feature Jacobians are the analytic ones of env_def.py:278-298,
    # 541-570, 700-761 written generically from the feature structure.
    def _obs_jac(self, z_fn, x, dim_in):
        z = z_fn(x)
        H = np.zeros(x.shape[:-1] + (z.shape[-1], dim_in))
        eye = np.eye(dim_in)
        for a in range(z.shape[-1]):
            H[..., a, :] = self._feature_row(a, x, dim_in, eye)
        return z, H

    def observe_jac(self, xu):
        """z, H = [C D] with z ~ z(mu) + H (xu - mu)."""
        return self._obs_jac(self.observe, xu, self.dim_xu)

    def observe_terminal_jac(self, x):
        return self._obs_jac(self.observe_terminal, x, self.dim_x)

    def forward_jac(self, xu, h=1e-30):
        f0 = self.dynamics(xu)
        J = np.zeros(xu.shape[:-1] + (self.dim_x, self.dim_xu))
        for i in range(self.dim_xu):
            d = np.zeros(self.dim_xu, dtype=complex)
            d[i] = 1j * h
            J[..., :, i] = np.imag(self.dynamics(xu + d)) / h
        return f0, J


class Linear(EnvSpec):
    """LinearDef + LinearBase: env_def.py:139-191, model.py:226-242."""

    name = "LinearKnown"
    dim_x, dim_u, dim_z, dim_z_term = 2, 1, 3, 2

    def __init__(self, A=None, B=None, xg=None, a=None, sig_x0=None, sig_eta=None, x0=None):
        self.x0 = np.array([5.0, 5.0]) if x0 is None else np.asarray(x0, float)
        self.xg = np.array([1.0, -1.0]) if xg is None else np.asarray(xg, float)
        self.A = np.array([[1.1, 0.0], [0.1, 1.1]]) if A is None else np.asarray(A, float)
        self.B = np.array([[0.1], [0.0]]) if B is None else np.asarray(B, float)
        # a = xg - A xg (env_def.py:167); batched A allowed
        self.a = (self.xg - np.einsum("...ij,...j->...i", self.A, self.xg)) if a is None else np.asarray(a, float)
        self.sig_x0 = 1e-20 * np.eye(2) if sig_x0 is None else sig_x0
        self.sig_eta = 1e-20 * np.eye(2) if sig_eta is None else sig_eta
        self.zg = np.concatenate((self.xg, np.zeros(self.xg.shape[:-1] + (1,))), axis=-1)
        self.zg_term = self.xg
        self.C = np.vstack((np.eye(2), np.zeros((1, 2))))
        self.D = np.array([[0.0], [0.0], [1.0]])
        self.c = np.zeros(3)

    def dynamics(self, xu):  # model.py:230-231   xu @ AB.T + a.T
        A, B, a = self.A, self.B, self.a
        x, u = xu[..., :2], xu[..., 2:]
        if A.ndim == 2:
            AB = np.concatenate((A, B), axis=1)
            return xu @ AB.T + a
        # per-problem A: A [Bt,2,2], xu [Bt,P,3] or [Bt,3]
        AB = np.concatenate((A, np.broadcast_to(B, A.shape[:-2] + B.shape[-2:])), axis=-1)
        if xu.ndim == AB.ndim:  # [Bt,P,n]
            return np.einsum("bpj,bij->bpi", xu, AB) + a[:, None, :]
        return np.einsum("bj,bij->bi", xu, AB) + a

    def observe(self, xu):  # env_def.py:175-177
        return xu + self.c

    def observe_terminal(self, x):  # env_def.py:183-185
        return x

    # --- linearisations (env_def.py:179-191, model.py:240-242)
    def observe_linearize(self, xu):
        return self.observe(xu), self.C, self.c, self.D

    def observe_terminal_linearize(self, x):
        return x, np.eye(2), np.zeros(2)

    def forward_linearize(self, xu):
        return self.dynamics(xu), self.A, self.B, self.a, self.sig_eta


class LinearMinimumEnergy(Linear):
    """LinearMinimumEnergyDef: env_def.py:194-230."""

    name = "LinearKnownMinimumEnergy"
    dim_x, dim_u, dim_z, dim_z_term = 2, 1, 1, 2

    def __init__(self):
        g = np.array([-5.0, -5.0])
        A = np.array([[1.05, 0.0], [0.05, 1.01]])
        super().__init__(A=A, B=np.array([[0.1], [0.0]]), xg=g, a=g - A @ g,
                         sig_x0=np.diag([1e-1, 5e0]), sig_eta=np.diag([1e-1, 1e-2]))
        self.zg = np.zeros(1)  # xag None -> zeros(dim_u)  (env_def.py:66-70)
        self.zg_term = g
        self.C = np.zeros((1, 2))
        self.D = np.eye(1)
        self.c = np.zeros(1)

    def observe(self, xu):  # env_def.py:219-220
        return xu[..., 2:]

    def observe_linearize(self, xu):
        return self.observe(xu), self.C, self.c, self.D


class Pendulum(EnvSpec):
    """PendulumKnown: env_def.py:233-309, env_autograd.py:5-19."""

    name = "PendulumKnown"
    dim_x, dim_u, dim_z, dim_z_term = 2, 1, 4, 3

    def __init__(self):
        self.x0 = np.array([np.pi, 0.0])
        self.sig_x0 = 1e-5 * np.eye(2)
        self.sig_eta = np.diag([1e-5, 1e-5])
        self.zg = np.array([0.0, 1.0, 0.0, 0.0])
        self.zg_term = np.array([0.0, 1.0, 0.0])

    @staticmethod
    def dynamics(xu):
        dt, m, l, d, g, u_mx = 0.05, 1.0, 1.0, 1e-2, 9.80665, 2.0
        th, thd = xu[..., 0], xu[..., 1]
        u = np.clip(xu[..., 2], -u_mx, u_mx)
        th_dot_dot = -3.0 * g / (2 * l) * np.sin(th + np.pi) - d * thd
        th_dot_dot = th_dot_dot + 3.0 / (m * l ** 2) * u
        x_dot = thd + th_dot_dot * dt
        x_pos = th + x_dot * dt
        return np.stack((x_pos, x_dot), axis=-1)

    def _feature_row(self, a, x, dim_in, eye):
        # z = [sin th, cos th, thd, (u)]
        if a == 0:
            return np.cos(x[..., 0])[..., None] * eye[0]
        if a == 1:
            return -np.sin(x[..., 0])[..., None] * eye[0]
        return np.broadcast_to(eye[a - 1], x.shape[:-1] + (dim_in,))

    @staticmethod
    def observe(xu):  # env_def.py:273-276
        return np.stack((np.sin(xu[..., 0]), np.cos(xu[..., 0]), xu[..., 1], xu[..., 2]), axis=-1)

    @staticmethod
    def observe_terminal(x):  # env_def.py:288-291
        return np.stack((np.sin(x[..., 0]), np.cos(x[..., 0]), x[..., 1]), axis=-1)


class PendulumActReg(Pendulum):
    """PendulumKnownActReg: env_def.py:312-346 (cost on u only, no terminal features)."""

    name = "PendulumKnownActReg"
    dim_x, dim_u, dim_z, dim_z_term = 2, 1, 1, 1
    has_terminal_obs = False

    def __init__(self):
        super().__init__()
        self.zg = np.zeros(1)
        self.zg_term = np.zeros(1)

    @staticmethod
    def observe(xu):
        return xu[..., 2:]

    @staticmethod
    def observe_terminal(x):
        return None


class Cartpole(EnvSpec):
    """CartpoleKnown: env_def.py:491-612, env_autograd.py:25-54."""

    name = "CartpoleKnown"
    dim_x, dim_u, dim_z, dim_z_term = 4, 1, 6, 5

    def __init__(self):
        self.x0 = np.array([0.0, np.pi, 0.0, 0.0])
        self.sig_x0 = 1e-5 * np.eye(4)
        self.sig_eta = np.diag([1e-8] * 4)
        self.zg = np.array([0.0, 0.0, 1.0, 0.0, 0.0, 0.0])
        self.zg_term = np.array([0.0, 0.0, 1.0, 0.0, 0.0])

    @staticmethod
    def dynamics(xu):
        g, Mc, Mp, l = 9.81, 0.37, 0.127, 0.3365
        Mt = Mc + Mp
        dt = 1 / 250.0
        _u = np.clip(xu[..., 4], -5.0, 5.0)
        th = xu[..., 1]
        dth2 = np.power(xu[..., 3], 2)
        sth, cth = np.sin(th), np.cos(th)
        _num = -Mp * l * sth * cth * dth2 + Mt * g * sth - _u * cth
        _denom = l * ((4.0 / 3.0) * Mt - Mp * cth ** 2)
        th_acc = _num / _denom
        x_acc = (Mp * l * sth * dth2 - Mp * l * th_acc * cth + _u) / Mt
        return np.stack(
            (xu[..., 0] + dt * xu[..., 2], xu[..., 1] + dt * xu[..., 3],
             xu[..., 2] + dt * x_acc, xu[..., 3] + dt * th_acc), axis=-1)

    def _feature_row(self, a, x, dim_in, eye):
        # z = [x, sin th, cos th, xd, thd, (u)]
        if a == 0:
            return np.broadcast_to(eye[0], x.shape[:-1] + (dim_in,))
        if a == 1:
            return np.cos(x[..., 1])[..., None] * eye[1]
        if a == 2:
            return -np.sin(x[..., 1])[..., None] * eye[1]
        return np.broadcast_to(eye[a - 1], x.shape[:-1] + (dim_in,))

    @staticmethod
    def observe(xu):
        return np.stack((xu[..., 0], np.sin(xu[..., 1]), np.cos(xu[..., 1]),
                         xu[..., 2], xu[..., 3], xu[..., 4]), axis=-1)

    @staticmethod
    def observe_terminal(x):
        return np.stack((x[..., 0], np.sin(x[..., 1]), np.cos(x[..., 1]), x[..., 2], x[..., 3]), axis=-1)


class DoubleCartpole(EnvSpec):
    """DoubleCartpoleKnown: env_def.py:615-761, env_autograd.py:60-167."""

    name = "DoubleCartpoleKnown"
    dim_x, dim_u, dim_z, dim_z_term = 6, 1, 9, 8

    def __init__(self):
        self.x0 = np.array([0.0, np.pi, np.pi, 0.0, 0.0, 0.0])
        self.sig_x0 = 1e-6 * np.eye(6)
        self.sig_eta = np.diag([1e-6] * 6)
        self.zg = np.array([0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0])
        self.zg_term = self.zg[:8]

    @staticmethod
    def dynamics(xu):
        dt = 1 / 125
        g, Mc, Mp1, Mp2 = 9.81, 0.37, 0.127, 0.127
        Mt = Mc + Mp1 + Mp2
        L1 = L2 = 0.3365
        l1, l2 = L1 / 2, L2 / 2
        J1, J2 = Mp1 * L1 / 12, Mp2 * L2 / 12
        u_mx, input_amp = 10.0, 3.0
        th1, th2 = xu[..., 1], xu[..., 2]
        th_dot1, th_dot2 = xu[..., 4], xu[..., 5]
        sth1, cth1, sth2, cth2 = np.sin(th1), np.cos(th1), np.sin(th2), np.cos(th2)
        sdth, cdth = np.sin(th1 - th2), np.cos(th1 - th2)
        l1_mp1_mp2 = Mp1 * l1 + Mp2 * L2
        l1_mp1_mp2_cth1 = l1_mp1_mp2 * cth1
        Mp2_l2 = Mp2 * l2
        Mp2_l2_cth2 = Mp2_l2 * cth2
        l1_l2_Mp2 = L1 * l2 * Mp2
        l1_l2_Mp2_cdth = l1_l2_Mp2 * cdth
        one = np.ones_like(th1)
        zero = np.zeros_like(th1)
        M = np.stack((
            np.stack((Mt * one, l1_mp1_mp2_cth1, Mp2_l2_cth2), axis=-1),
            np.stack((l1_mp1_mp2_cth1, ((l1 ** 2) * Mp1 + (L1 ** 2) * Mp2 + J1) * one, l1_l2_Mp2_cdth), axis=-1),
            np.stack((Mp2_l2_cth2, l1_l2_Mp2_cdth, ((l2 ** 2) * Mp2 + J2) * one), axis=-1),
        ), axis=-2)
        C12 = -l1_mp1_mp2 * th_dot1 * sth1
        C13 = -Mp2_l2 * th_dot2 * sth2
        C23 = l1_l2_Mp2 * th_dot2 * sdth
        C32 = -l1_l2_Mp2 * th_dot1 * sdth
        C = np.stack((
            np.stack((zero, C12, C13), axis=-1),
            np.stack((zero, zero, C23), axis=-1),
            np.stack((zero, C32, zero), axis=-1),
        ), axis=-2)
        G = np.stack((zero, -(Mp1 * l1 + Mp2 * L1) * g * sth1, -Mp2 * l2 * g * sth2), axis=-1)
        u = input_amp * np.clip(xu[..., 6], -u_mx, u_mx)
        action = np.stack((u, zero, zero), axis=-1)
        M_inv = np.linalg.inv(M)
        C_x_dot = np.einsum("...ij,...j->...i", C, xu[..., 3:6])
        x_dot_dot = np.einsum("...ij,...j->...i", M_inv, action - C_x_dot - G)
        x_dot = xu[..., 3:6] + x_dot_dot * dt
        x_pos = xu[..., :3] + x_dot * dt
        return np.concatenate((x_pos, x_dot), axis=-1)

    def _feature_row(self, a, x, dim_in, eye):
        # z = [x, s1, c1, s2, c2, xd, thd1, thd2, (u)]
        if a == 0:
            return np.broadcast_to(eye[0], x.shape[:-1] + (dim_in,))
        if a in (1, 3):
            ang = 1 if a == 1 else 2
            return np.cos(x[..., ang])[..., None] * eye[ang]
        if a in (2, 4):
            ang = 1 if a == 2 else 2
            return -np.sin(x[..., ang])[..., None] * eye[ang]
        return np.broadcast_to(eye[a - 2], x.shape[:-1] + (dim_in,))

    @staticmethod
    def observe(xu):
        return np.stack((xu[..., 0], np.sin(xu[..., 1]), np.cos(xu[..., 1]), np.sin(xu[..., 2]),
                         np.cos(xu[..., 2]), xu[..., 3], xu[..., 4], xu[..., 5], xu[..., 6]), axis=-1)

    @staticmethod
    def observe_terminal(x):
        return np.stack((x[..., 0], np.sin(x[..., 1]), np.cos(x[..., 1]), np.sin(x[..., 2]),
                         np.cos(x[..., 2]), x[..., 3], x[..., 4], x[..., 5]), axis=-1)


# ---------------------------------------------------------------- quadrotor
QUAD_W = 600 / 30.0
QUAD_H = 400 / 30.0
QUAD_DX = QUAD_W / 25  # vehicle_dx  (mpc_quad.py:73)
QUAD_DY = QUAD_H / 100  # vehicle_dy
QUAD_MASS = 5.0 * (2 * QUAD_DX) * (2 * QUAD_DY)  # density * area (mpc_quad.py:286)
QUAD_INERTIA = QUAD_MASS * ((2 * QUAD_DX) ** 2 + (2 * QUAD_DY) ** 2) / 12.0
QUAD_FS = 10


class Quadrotor(EnvSpec):
    """fp64 planar-quadrotor RESTATEMENT of QuadrotorDef (mpc_quad.py:219-383).

    PARITY UNPINNED for ``dynamics``: the reference steps a Box2D (float32, un-vendored,
    unpinned) world per sigma point (mpc_quad.py:325-360).  Box2D is absent from this
    image; this is the published semi-implicit Euler update Box2D's island solver applies
    to a single free body (v += h(g + F/m); w += h tau/I; w *= 1/(1+h*angularDamping);
    p += h v; th += h w), without the per-step translation/rotation clamps and without
    wall contacts.  The same function is monkey-patched into the reference
    (tests/golden/make_golden.py) so 1e-9 parity is judged between two implementations of
    the SAME fp64 dynamics.  ``measure`` is pure NumPy in the reference and is restated
    including its typos (mpc_quad.py:371-383).
    """

    name = "Quadrotor"
    dim_x, dim_u, dim_z, dim_z_term, dim_y = 6, 2, 8, 6, 8

    def __init__(self, sig_zeta=None):
        self.x0 = np.array([QUAD_W / 4, QUAD_H / 2, 0.0, 0.0, 0.0, 0.0])
        self.sig_x0 = 1e-5 * np.eye(6)
        self.sig_eta = np.diag([1e-6] * 2 + [1e-6] + [1e-4] * 2 + [1e-4])
        xag = np.array([3 * QUAD_W / 4, QUAD_H / 2, 0.0, 0.0, 0.0, 0.0])
        self.zg = np.concatenate((xag, np.zeros(2)))
        self.zg_term = xag
        self.sig_zeta = sig_zeta
        self.gravity = 9.81 * QUAD_MASS  # mpc_quad.py:321-323

    @staticmethod
    def dynamics(xu):
        h = 1.0 / QUAD_FS
        u1 = np.clip(xu[..., 6], 0.0, 30.0)
        u2 = np.clip(xu[..., 7], 0.0, 30.0)
        psi = xu[..., 2]
        s, c = np.sin(psi), np.cos(psi)
        f = u1 + u2
        vx = xu[..., 3] + h * ((-s * f) / QUAD_MASS)
        vy = xu[..., 4] + h * (-9.81 + (c * f) / QUAD_MASS)
        w = xu[..., 5] + h * (QUAD_DX * (u2 - u1)) / QUAD_INERTIA
        w = w * (1.0 / (1.0 + h * 0.5))
        return np.stack((xu[..., 0] + h * vx, xu[..., 1] + h * vy, psi + h * w, vx, vy, w), axis=-1)

    @staticmethod
    def observe(xu):
        return xu

    @staticmethod
    def observe_terminal(x):
        return x

    @staticmethod
    def measure(x):
        dx = QUAD_DX
        c, s = np.cos(x[..., 2]), np.sin(x[..., 2])
        lx = x[..., 0] - dx * c
        ly = x[..., 1] - dx * s
        lxd = x[..., 3] - dx * -s * x[..., 5]
        lyd = x[..., 4] - dx * c * x[..., 5]
        rx = x[..., 0] + dx * c
        ry = x[..., 1] + dx * s
        rxd = x[..., 3] + dx - s * x[..., 5]
        ryd = x[..., 4] + dx + c * x[..., 5]
        return np.stack((lx, ly, rx, ry, lxd, lyd, rxd, ryd), axis=-1)


REGISTRY = {
    "LinearKnown": Linear,
    "LinearKnownMinimumEnergy": LinearMinimumEnergy,
    "PendulumKnown": Pendulum,
    "PendulumKnownActReg": PendulumActReg,
    "CartpoleKnown": Cartpole,
    "DoubleCartpoleKnown": DoubleCartpole,
    "Quadrotor": Quadrotor,
}


def make(name, **kw):
    return REGISTRY[name](**kw)
