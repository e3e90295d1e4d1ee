"""Model registry mirror (reference: i2c/model.py:19-44, i2c/env_def.py).  A model carries the constants i2c reads
from ``sys`` and *markers* for the environment maps: dynamics and cost features are evaluated inside the CUDA
kernels (csrc/envs.cuh), so ``sys.observe`` / ``sys.forward`` are only meaningful as arguments of
``QuadratureInference`` or ``I2cGraph``."""
import numpy as np

from i2c_b200 import envs as _envs


# axis labels / units of the state-action and cost-feature components: read by the scripts' own plotting code
# (e.g. nonlinear_covariance_control.py:63) -- data of the definitions, env_def.py:142-144, 237-239, 495-504, 619-639
_LABELS = {
    "LinearKnown": (["$x_1$", "$x_2$", "$u$"], None, [None] * 3),
    "LinearKnownMinimumEnergy": (["$x_1$", "$x_2$", "$u$"], ["$u$"], [None] * 3),
    "PendulumKnown": (["$\\theta$", "$\\dot{\\theta}$", "$u$"],
                      ["$\\sin(\\theta)$", "$\\cos(\\theta)$", "$\\dot{\\theta}$", "$u$"], ["rad", "rad/s", "Nm"]),
    "PendulumKnownActReg": (["$\\theta$", "$\\dot{\\theta}$", "$u$"], ["$u$"], ["rad", "rad/s", "Nm"]),
    "CartpoleKnown": (["$x$", "$\\theta$", "$\\dot{x}$", "$\\dot{\\theta}$", "$u$"],
                      ["$x$", "$\\sin(\\theta)$", "$\\cos(\\theta)$", "$\\dot{x}$", "$\\dot{\\theta}$", "$u$"],
                      ["m", "rad", "m/s", "rad/s", "Nm"]),
    "DoubleCartpoleKnown": (["$x$", "$\\theta_1$", "$\\theta_2$", "$\\dot{x}$", "$\\dot{\\theta}_1$", "$\\dot{\\theta}_2$", "$u$"],
                            ["$x$", "$\\sin\\theta_1$", "$\\cos\\theta_1$", "$\\sin\\theta_2$", "$\\cos\\theta_2$", "$\\dot{x}$",
                             "$\\dot{\\theta}_1$", "$\\dot{\\theta}_2$", "$u$"], ["m", "rad", "rad", "m/s", "rad/s", "rad/s", "Nm"]),
    "Quadrotor": (["$x$", "$y$", "$\\psi$", "$\\dot{x}$", "$\\dot{y}$", "$\\dot{\\psi}$", "$u_1$", "$u_2$"], None, [None] * 8),
}


class KnownModel(object):
    data_driven = False
    model = None

    def __init__(self, env_name):
        key, z_key, unit = _LABELS[env_name]
        self.key, self.z_key, self.unit = list(key), list(z_key if z_key is not None else key), list(unit)
        c = _envs.make(env_name)
        self._b200_env = env_name
        self.name = env_name
        self.dim_x, self.dim_u, self.dim_z, self.dim_z_term = c.dim_x, c.dim_u, c.dim_z, c.dim_z_term
        self.dim_y = c.dim_y
        self.x0 = c.x0.reshape(-1, 1).copy()
        self.sig_x0 = c.sig_x0.copy()
        self.sig_eta = c.sig_eta.copy()
        # zg follows xag like the reference's BaseDef._zg (env_def.py:66-70): scripts overwrite model.xag
        self.xag = c.zg.reshape(-1, 1)[: self.dim_z - self.dim_u].copy() if self.dim_z > self.dim_u else None
        self.zg_term = c.zg_term.reshape(-1, 1).copy()
        self.has_terminal_features = c.has_term
        self.u_lim = c.u_lim
        self.sig_zeta = None
        for k in ("A", "B", "a"):
            if hasattr(c, k):
                setattr(self, k, np.array(getattr(c, k), float).reshape((-1, 1)) if k == "a" else np.array(getattr(c, k), float))

    @property
    def zg(self):
        if self.xag is not None:
            return np.vstack((np.asarray(self.xag, float).reshape(-1, 1), np.zeros((self.dim_u, 1))))
        return np.zeros((self.dim_u, 1))

    # dims ---------------------------------------------------------------------------------------
    @property
    def dim_xu(self):
        return self.dim_x + self.dim_u

    dim_s = dim_xu

    def _b200_env_par(self):
        if hasattr(self, "A"):
            return _envs.linear_params(self.A, self.B, np.asarray(self.a, float).reshape(-1))
        return None

    def clip_u(self, u):
        lo, hi = self.u_lim
        return np.clip(u, lo, hi)

    def init(self):
        return self.x0.squeeze(), self.sig_x0

    # markers -------------------------------------------------------------------------------------
    def _in_kernel(self, *a, **k):
        raise NotImplementedError("environment maps are evaluated inside the CUDA kernels; pass this method to "
                                  "QuadratureInference / I2cGraph instead of calling it (no CPU fallback)")

    def observe(self, xu):
        return self._in_kernel()

    def observe_terminal(self, x):
        return self._in_kernel()

    def observe_terminal_x(self, x):
        return self._in_kernel()

    def forward(self, xu):
        return self._in_kernel()

    def dynamics(self, xu):
        return self._in_kernel()

    def measure(self, x):
        return self._in_kernel()


class BaseModel(object):
    """Interface of the reference's model base class (i2c/model.py:47-150) as far as the known-model path uses it."""

    model = None
    data_driven = False

    def init(self):
        return np.asarray(self.x0, float).squeeze(), self.sig_x0

    def observe_terminal_x(self, x):
        return self.observe_terminal(x)

    def calibrate_epistemic(self, y):
        assert self.model is None, self.model

    def save(self, path):
        print("Known model, no saving")


class BaseModelKnown(BaseModel):
    """Base of the known models (i2c/model.py:146-175): mixed with an environment definition, e.g.
    ``class QuadrotorKnown(QuadrotorDef, BaseModelKnown)`` (mpc_quad.py:386).  I2cGraph, QuadratureInference and the MPC
    policies take the constants from the definition part and run the maps of the matching in-kernel environment
    (env_def.BaseDef._b200_env); ``forward`` exists for API compatibility and evaluates the definition's own NumPy
    ``dynamics`` if it has one (host-side simulation in the scripts), never on the inference path."""

    data_driven = False

    def __init__(self, model=None, model_def=None):
        assert model is None
        assert model_def is None
        super().__init__()

    def forward(self, xu):
        _x = self.dynamics(xu)
        return _x, np.repeat(self.sig_eta[None, :, :], xu.shape[0], axis=0)

    def predict(self, xu):
        return self.dynamics(xu)


_LOOKUP = {  # i2c/model.py:25-36
    "LinearKnown": "LinearKnown",
    "LinearKnownMinimumEnergy": "LinearKnownMinimumEnergy",
    "PendulumKnown": "PendulumKnown",
    "PendulumKnownActReg": "PendulumKnownActReg",
    "CartpoleKnown": "CartpoleKnown",
    "DoubleCartpoleKnown": "DoubleCartpoleKnown",
    "Quadrotor": "Quadrotor",
}


def make_env_model(env_def, model_def=None):
    if model_def is not None:
        raise NotImplementedError("learned models are not part of the CUDA path (never instantiated upstream either)")
    if env_def not in _LOOKUP:
        raise KeyError(f"{env_def!r} is not registered with the CUDA path (no CPU fallback)")
    return KnownModel(_LOOKUP[env_def])


def QuadrotorKnown():
    """scripts/mpc_state_est/mpc_quad.py:386 (fp64 restatement of the Box2D step)."""
    return KnownModel("Quadrotor")
