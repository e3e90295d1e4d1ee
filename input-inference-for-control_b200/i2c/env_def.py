"""Environment definitions (mirror of the interface of the reference's i2c/env_def.py).

The reference's definitions carry two things: constants that i2c reads from ``sys`` (dimensions, start-state belief,
process noise, goal, action limits) and the NumPy dynamics / cost-feature maps.  On the CUDA path the maps live in the
kernels (csrc/envs.cuh); what remains on the host is the constant surface and the helper properties scripts rely on
(`BaseDef`: i2c/env_def.py:12-136).  A script may define its own subclass -- scripts/mpc_state_est/mpc_quad.py:219,386
does: ``class QuadrotorDef(BaseDef)``, ``class QuadrotorKnown(QuadrotorDef, BaseModelKnown)`` -- the model is then matched to
its in-kernel counterpart by ``name`` (DEVICE_ENVS); an unknown name raises: there is no CPU fallback for the maps.
"""
import numpy as np

from i2c_b200 import envs as _envs

# definition ``name`` attribute (or class name) -> environment compiled into the CUDA library (include/i2c_b200.h: i2c_env)
DEVICE_ENVS = {
    "2D Quadrator": "Quadrotor",  # [sic] mpc_quad.py:222
    "QuadrotorDef": "Quadrotor",
    "QuadrotorKnown": "Quadrotor",
}
for _n in ("LinearKnown", "LinearKnownMinimumEnergy", "PendulumKnown", "PendulumKnownActReg", "CartpoleKnown",
           "DoubleCartpoleKnown", "Quadrotor"):
    DEVICE_ENVS[_n] = _n


class BaseDef(object):
    """Base definition with the helpers of the reference (i2c/env_def.py:12-136)."""

    def __init__(self, *args, **kwargs):
        self._pickle_args, self._pickle_kwargs = args, kwargs
        super().__init__(*args, **kwargs)

    def __getstate__(self):
        return {"_pickle_args": self._pickle_args, "_pickle_kwargs": self._pickle_kwargs}

    def __setstate__(self, d):
        out = type(self)(*d["_pickle_args"], **d["_pickle_kwargs"])
        self.__dict__.update(out.__dict__)

    name = "Template"
    deterministic = False
    dim_x = dim_xf = dim_u = None
    xag = x0 = x0_dist = xu_lim = None

    @property
    def random_starting_state(self):
        return self.x0_dist is not None

    @property
    def dim_s(self):
        return self.dim_x + self.dim_u

    dim_xu = dim_s

    @property
    def dim_xt(self):
        return self.dim_xf + self.dim_u

    @property
    def dim_yt(self):
        return self.dim_x

    @property
    def dim_xat(self):
        return self.dim_x + self.dim_u

    def _zg(self):
        if self.xag is not None:
            return np.vstack((self.xag, np.zeros((self.dim_u, 1))))
        return np.zeros((self.dim_u, 1))

    @property
    def zg(self):
        return self._zg()

    @property
    def zg_term(self):
        return self.zg

    @property
    def zgc(self):
        return np.vstack((self.xg, np.zeros((self.dim_u, 1))))

    def observe(self, xu):
        raise NotImplementedError

    def observe_linearize(self, xu):
        raise NotImplementedError

    def observe_terminal(self, x):
        raise NotImplementedError

    def observe_terminal_linearize(self, x):
        raise NotImplementedError

    def remove_state_bounds(self):
        pass

    def xu_in_bounds(self, xu):
        lo, hi = self.xu_lim[0, :, None], self.xu_lim[1, :, None]
        return bool(np.all(lo < xu) and np.all(hi > xu))

    def x_in_bounds(self, xu):
        x = xu[:, : self.dim_x]
        lo, hi = self.xu_lim[0, : self.dim_x, None], self.xu_lim[1, : self.dim_x, None]
        return bool(np.all(lo < x) and np.all(hi > x))

    def clip_u(self, u):
        return np.clip(u, self.xu_lim[0, self.dim_x:], self.xu_lim[1, self.dim_x:])

    # ---- CUDA path ----------------------------------------------------------------------------------
    @property
    def _b200_env(self):
        """Name of the in-kernel environment this definition stands for."""
        for key in (getattr(self, "name", None), type(self).__name__, *[c.__name__ for c in type(self).__mro__]):
            if key in DEVICE_ENVS:
                return DEVICE_ENVS[key]
        raise KeyError(f"environment definition {type(self).__name__!r} (name {getattr(self, 'name', None)!r}) has no in-kernel "
                       "counterpart: the dynamics / cost-feature maps of the CUDA path are compiled (csrc/envs.cuh); there is "
                       "no CPU fallback")

    def _b200_env_par(self):
        if all(hasattr(self, k) for k in ("A", "B", "a")) and self._b200_env.startswith("Linear"):
            return _envs.linear_params(self.A, self.B, np.asarray(self.a, float).reshape(-1))
        return None


def _constants(env_name):
    c = _envs.make(env_name)
    d = dict(name=env_name, dim_x=c.dim_x, dim_u=c.dim_u, dim_z=c.dim_z, dim_z_term=c.dim_z_term, dim_y=c.dim_y,
             x0=c.x0.reshape(-1, 1).copy(), sig_x0=c.sig_x0.copy(), sig_eta=c.sig_eta.copy(),
             xag=c.zg.reshape(-1, 1)[: c.dim_z - c.dim_u].copy() if c.dim_z > c.dim_u else None)
    if c.u_lim is not None:
        lo, hi = c.u_lim
        lo, hi = np.broadcast_to(lo, (c.dim_u,)), np.broadcast_to(hi, (c.dim_u,))
    else:  # linear systems: unbounded actions
        lo, hi = np.full(c.dim_u, -np.inf), np.full(c.dim_u, np.inf)
    d["xu_lim"] = np.array([[-np.inf] * c.dim_x + list(lo), [np.inf] * c.dim_x + list(hi)])
    return d


def _make_def(cls_name, env_name, doc):
    body = _constants(env_name)
    body["__doc__"] = doc
    zt = _envs.make(env_name).zg_term.reshape(-1, 1).copy()
    body["zg_term"] = property(lambda self, _zt=zt: _zt)
    return type(cls_name, (BaseDef,), body)


LinearDef = _make_def("LinearDef", "LinearKnown", "LDS of the LQR-equivalence experiment (env_def.py:139-191).")
LinearMinimumEnergyDef = _make_def("LinearMinimumEnergyDef", "LinearKnownMinimumEnergy", "env_def.py:194-230.")
PendulumKnown = _make_def("PendulumKnown", "PendulumKnown", "env_def.py:233-309.")
PendulumKnownActReg = _make_def("PendulumKnownActReg", "PendulumKnownActReg", "env_def.py:312-346.")
CartpoleKnown = _make_def("CartpoleKnown", "CartpoleKnown", "env_def.py:500-570.")
DoubleCartpoleKnown = _make_def("DoubleCartpoleKnown", "DoubleCartpoleKnown", "env_def.py:634-732.")
