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
    """Declared for import compatibility; degree**dim point rules are not built for the CUDA path."""

    degree: int

    def __post_init__(self):
        raise NotImplementedError("Gauss-Hermite quadrature is not available on the CUDA path (no CPU fallback)")
