"""Experiment hyper-parameter types (mirror of the reference's i2c/exp_types.py:9-68)."""
from dataclasses import dataclass

import numpy as np


@dataclass
class GaussianI2c(object):
    inference: object
    alpha: float
    alpha_update_tol: float
    Q: np.ndarray
    Qf: np.ndarray
    R: np.ndarray
    mu_u: np.ndarray
    sig_u: np.ndarray
    mu_x_term: np.ndarray
    sig_x_term: np.ndarray


@dataclass
class Linearize(object):
    """Linearisation-based inference (no parameters)."""


@dataclass
class CubatureQuadrature(object):
    """Spherical-radial cubature / unscented rule: points [0; +I; -I] (exp_types.py:30-49)."""

    alpha: float
    beta: float
    kappa: float

    @staticmethod
    def pts(dim):
        eye = np.eye(dim)
        return np.vstack((np.zeros((1, dim)), eye, -eye))

    def weights(self, dim):
        if not self.alpha > 0:
            raise AssertionError("alpha must be positive")
        lam = self.alpha ** 2 * (dim + self.kappa) - dim
        scale = np.sqrt(dim + lam)
        w_mu = np.full(2 * dim + 1, 1.0 / (2.0 * (dim + lam)))
        w_mu[0] *= 2.0 * lam
        w_sig = w_mu.copy()
        w_sig[0] += 1.0 - self.alpha ** 2 + self.beta
        return scale, w_mu, w_sig


@dataclass
class GaussHermiteQuadrature(object):
    """Tensor-product Gauss-Hermite rule with degree**dim points (exp_types.py:52-68).  The 1-D nodes / weights come
    from the CUDA library (i2c_gauss_hermite), the same rule its kernels integrate with."""

    degree: int

    def __post_init__(self):
        if self.degree < 1:
            raise AssertionError("degree must be >= 1")
        import i2c_b200

        x, w = i2c_b200.batched.gauss_hermite(self.degree)
        self.gh_pts, self.gh_weights = x, w * np.sqrt(np.pi)  # hermgauss convention: weights sum to sqrt(pi)

    def pts(self, dim):
        axes = np.meshgrid(*([self.gh_pts] * dim))
        return np.stack([a.ravel() for a in axes], axis=1)

    def weights(self, dim):
        axes = np.meshgrid(*([self.gh_weights] * dim))
        w = np.prod(np.stack([a.ravel() for a in axes], axis=1), axis=1) / np.pi ** (dim / 2)
        return np.sqrt(2), w, w
