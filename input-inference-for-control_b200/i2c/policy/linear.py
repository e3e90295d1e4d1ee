"""Time-indexed linear-Gaussian controller containers (mirror of i2c/policy/linear.py:9-90): pure host-side
consumers of the controllers extracted by the CUDA path."""
import numpy as np


class TimeIndexedLinearGaussianPolicy(object):
    def __init__(self, sig_u, H, dim_u, dim_x, control_step=1):
        self.H, self.dim_u, self.dim_x = H, dim_u, dim_x
        self.sig_u = np.asarray(sig_u, float)
        self.control_step = control_step
        self.init()

    def init(self):
        self.zero()
        self.sig_k = np.tile(self.sig_u[None], (self.H, 1, 1))

    def zero(self):
        self.K = np.zeros((self.H, self.dim_u, self.dim_x))
        self.k = np.zeros((self.H, self.dim_u))
        self.sig_k = np.zeros((self.H, self.dim_u, self.dim_u))

    def write(self, K, k, sig_k):
        self.K[...] = K
        self.k[...] = k
        self.sig_k[...] = sig_k

    def __call__(self, i, x, deterministic=True):
        assert i < self.H
        if i % self.control_step == 0:
            mean = self.K[i] @ x + self.k[i][:, None]
            if deterministic:
                self.u = mean
            else:
                self.u = np.random.multivariate_normal(mean[:, 0], self.sig_k[i], 1)
        return self.u


class ExpertTimeIndexedLinearGaussianPolicy(object):
    hard_exp_threshold = 3.0

    def __init__(self, sig_u, H, dim_u, dim_x, soft=True):
        self.H, self.dim_u, self.dim_x = H, dim_u, dim_x
        self.sig_u = np.asarray(sig_u, float)
        self.soft = soft
        self.init()

    def init(self):
        self.K = np.zeros((self.H, self.dim_u, self.dim_x))
        self.k = np.zeros((self.H, self.dim_u))
        self.sig_k = np.tile(self.sig_u[None], (self.H, 1, 1))
        self.mu = np.zeros((self.H, self.dim_x))
        self.lam = np.ones((self.H, self.dim_x, self.dim_x))

    zero = init

    def write(self, K, k, sig_k, mu, lam):
        self.K[...], self.k[...], self.sig_k[...], self.mu[...], self.lam[...] = K, k, sig_k, mu, lam

    def __call__(self, i, x, deterministic=True):
        assert i < self.H
        d = x - self.mu[i][:, None]
        e = 0.5 * (d.T @ self.lam[i] @ d).item()
        gate = np.exp(-e) if self.soft else float(abs(e) < self.hard_exp_threshold)
        mean = (self.k[i][:, None] + gate * (self.K[i] @ d)).reshape(self.dim_u, 1)
        if deterministic:
            return mean
        noise = np.random.multivariate_normal(np.zeros(self.dim_u), self.sig_k[i].reshape(self.dim_u, self.dim_u), 1)
        return mean + noise.reshape(self.dim_u, 1)
