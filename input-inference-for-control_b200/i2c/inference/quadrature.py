"""QuadratureInference mirror (reference: i2c/inference/quadrature.py:7-58) backed by the stand-alone
sigma-point kernel (``i2c_quadrature``).  ``f`` must be one of the registered environment maps of a model made by
``i2c.model.make_env_model`` (``sys.observe``, ``sys.observe_terminal_x``, ``sys.forward``, ``sys.measure``):
arbitrary Python callables cannot be evaluated in-kernel and raise (no CPU fallback)."""
import numpy as np

import i2c_b200
from i2c.exp_types import CubatureQuadrature, GaussHermiteQuadrature

_FN = {"observe": "observe", "observe_terminal": "observe_terminal", "observe_terminal_x": "observe_terminal",
       "forward": "forward", "dynamics": "forward", "measure": "measure"}


class QuadratureInference(object):
    def __init__(self, params, dim):
        if not isinstance(params, (CubatureQuadrature, GaussHermiteQuadrature)):  # quadrature.py:9
            raise AssertionError("params must be CubatureQuadrature or GaussHermiteQuadrature")
        self.params = params
        self.dim = dim
        self.base_pts = params.pts(dim)
        self.sf, self.weights_mu, self.weights_sig = params.weights(dim)
        self.n_points = self.base_pts.shape[0]

    def _resolve(self, f):
        model = getattr(f, "__self__", None)
        name = getattr(f, "__name__", "")
        env = getattr(model, "_b200_env", None)
        if env is None or name not in _FN:
            raise NotImplementedError(
                f"QuadratureInference on the CUDA path only evaluates registered environment maps, got {f!r}")
        return model, env, _FN[name]

    def _run(self, f, m_x, sig_x):
        model, env, fn = self._resolve(f)
        m = np.asarray(m_x, float).reshape(1, self.dim)
        if isinstance(self.params, GaussHermiteQuadrature):
            kw = dict(gh_degree=int(self.params.degree))
        else:
            kw = dict(quad=(self.params.alpha, self.params.beta, self.params.kappa))
        my, Sy, Sxy, st = i2c_b200.quadrature(env, fn, m, np.asarray(sig_x, float)[None], env_par=model._b200_env_par(),
                                              device=getattr(model, "device", 0), **kw)
        if st[0] != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")  # quadrature.py:17-24
        self.m_y, self.sig_y, self.sig_xy = my, Sy[0], Sxy[0]
        return model

    def forward(self, f, m_x, sig_x):
        self._run(f, m_x, sig_x)
        return self.m_y.T, self.sig_y

    def forward_gaussian(self, f, m_x, sig_x):
        model = self._run(f, m_x, sig_x)
        self.sig_noise = np.array(model.sig_eta, float)  # sum_p w_p Sigma_eta = Sigma_eta for known models
        return self.m_y.T, self.sig_y, self.sig_noise
