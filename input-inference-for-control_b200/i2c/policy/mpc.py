"""i2c MPC controllers (mirror of i2c/policy/mpc.py) on the CUDA path: the planning graph stays on the device,
the horizon shift is the O(1) ring rotation of ``i2c_shift_horizon`` instead of list pop/append + deepcopy."""
import copy

import numpy as np

import i2c_b200


class MpcPolicy(object):
    def __init__(self, i2c, n_iter, sig_u, z_traj=None):
        self.dim_u, self.dim_x = i2c.sys.dim_u, i2c.sys.dim_x
        self.sig_u = sig_u
        self.i2c = i2c
        i2c.tau = 0
        self.model = i2c.sys
        self.n_iter = n_iter
        self.z_traj = None if z_traj is None else np.asarray(z_traj, float)
        if self.z_traj is not None:
            for i, c in enumerate(i2c.cells):
                c.z = self.z_traj[i, :, None]
            self._z_last = self.z_traj[i2c.H - 1].copy()
        self.xu_history, self.z_history = [], []
        self._snapshot_init = None  # device snapshot taken lazily by reset() users (i2c_init = deepcopy(i2c) upstream)
        self.i2c_init = copy.deepcopy(i2c)  # policy/mpc.py:23

    def set_control(self, feedforward):
        self.i2c.tau = 0 if feedforward else self.i2c.H

    def reset(self):
        """policy/mpc.py:44-48: back to the graph the policy was built with."""
        self.i2c.close()
        self.i2c = copy.deepcopy(self.i2c_init)
        self.model = self.i2c.sys
        if self.z_traj is not None:
            self._z_last = self.z_traj[self.i2c.H - 1].copy()
        self.xu_history, self.z_history = [], []

    def update_models(self, sys):
        """policy/mpc.py:87-89 (the freshly appended cells take their constants from the graph's sys here)."""
        self.i2c.sys = sys
        self.model = sys

    def optimize(self, n_iter, x):
        assert x.shape == (self.dim_x, 1), f"{x.shape}, {(self.dim_x, 1)}"
        self.i2c.sys.x0 = x
        self.i2c._run(n_iter, i2c_b200.capi.PH_FORWARD | i2c_b200.capi.PH_BACKWARD | i2c_b200.capi.PH_UPDATE_PRIORS)

    def _shift(self, i):
        g = self.i2c
        if self.z_traj is not None:
            z_new = self.z_traj[i + g.H] if (i + g.H) < self.z_traj.shape[0] else self._z_last
            self._z_last = np.array(z_new, float)
        else:
            z_new = np.asarray(g.sys.zg, float).reshape(-1)
        g._push()
        g._g.shift_horizon(z_new, g._cell_init_mu_u, g.alpha_base)
        # mirror the rotation on the host-side views: cells.pop(0); cells.append(deepcopy(cell_init))
        old = g.cells.pop(0)
        new = type(old)(g, 0)
        object.__setattr__(new, "_z", np.array(z_new, float).reshape(-1, 1))
        g.cells.append(new)
        g._cache = {}

    def __call__(self, i, x, deterministic=True):
        self.optimize(self.n_iter, x)
        # policy/mpc.py:59.  In the reference this very call raises TypeError (compute_update_alpha takes no keyword
        # arguments, i2c.py:921: SURVEY.md "known-broken reference code"); the intended alpha adaptation is done here.
        self.i2c.compute_update_alpha(True, calc_evar=False, calc_propagate=False)
        self.xu_history.append(self.i2c.get_marginal_state_action())
        self.z_history.append(self.i2c.get_marginal_observed_trajectory()[0])
        mu, sig = self.i2c.cells[0].mu_u0_m.copy(), self.i2c.cells[0].sig_u0_m.copy()
        u = mu.reshape((self.dim_u, 1))
        if not deterministic:
            u = np.random.multivariate_normal(mu.squeeze(), sig, 1).reshape((self.dim_u, 1))
        self._shift(i)
        return u


class PartiallyObservedMpcPolicy(MpcPolicy):
    def __init__(self, i2c, n_iter, sig_u, z_traj=None):
        super().__init__(i2c, n_iter, sig_u, z_traj)
        self.mu = np.array(i2c.sys.x0, float)
        self.covar = np.array(i2c.sys.sig_x0, float)
        self.mus, self.covars = [], []

    def filter(self, y, u):
        assert u.shape == (self.dim_u, 1)
        g = self.i2c
        g.sys.x0, g.sys.sig_x0 = self.mu, self.covar
        g._push()
        g._g.ckf_step(np.asarray(y, float).reshape(1, -1), np.asarray(u, float).reshape(1, -1), g.sys.sig_zeta)
        m, c = g._g.get_initial_state()
        self.mu, self.covar = m[0].reshape(-1, 1), c[0]
        g.sys.x0, g.sys.sig_x0 = self.mu, self.covar
        g._x0_pushed = (np.asarray(self.mu, float).tobytes(), np.asarray(self.covar, float).tobytes())
        return self.mu, self.covar

    def optimize(self, n_iter, mu, covar):
        assert mu.shape == (self.dim_x, 1), f"{mu.shape}, {(self.dim_x, 1)}"
        self.i2c.sys.x0 = mu
        self.i2c.sys.sig_x0 = covar
        self.i2c._run(n_iter, i2c_b200.capi.PH_FORWARD | i2c_b200.capi.PH_BACKWARD | i2c_b200.capi.PH_UPDATE_PRIORS)

    def __call__(self, i, y, u, deterministic=True):
        if i > 0:
            self.filter(y, u)
        self.mus.append(self.mu)
        self.covars.append(self.covar)
        self.optimize(self.n_iter, self.mu, self.covar)
        self.xu_history.append(self.i2c.get_marginal_state_action())
        self.z_history.append(self.i2c.get_marginal_observed_trajectory()[0])
        mu, sig = self.i2c.cells[0].mu_u0_m.copy(), self.i2c.cells[0].sig_u0_m.copy()
        ctrl = mu.reshape((self.dim_u, 1))
        if not deterministic:
            ctrl = np.random.multivariate_normal(mu.squeeze(), sig, 1).reshape((self.dim_u, 1))
        self._shift(i)
        return ctrl
