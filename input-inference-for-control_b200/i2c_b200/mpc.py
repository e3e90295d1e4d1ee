"""Batched receding-horizon i2c controllers on the GPU.

``BatchedPartiallyObservedMpc`` is the batched counterpart of the reference's
``PartiallyObservedMpcPolicy`` (i2c/policy/mpc.py:115-182): B closed-loop roll-outs share one device-resident
planning graph (``BatchedI2c`` with horizon T_plan); every control step runs the cubature Kalman filter
(i2c_ckf_step), ``n_iter`` x (forward + backward sweep, _update_priors) in one persistent kernel launch, reads the
first action, and shifts the horizon in O(1) (ring buffer of cells, i2c_shift_horizon).
"""
import numpy as np

from . import _capi as capi


class BatchedPartiallyObservedMpc:
    def __init__(self, i2c, n_iter, sig_u, z_traj=None, sig_zeta=None, pinned_io=False):
        """pinned_io=True: the actions returned by the fused ``__call__`` are views into a ring of three page-locked host
        buffers (each stays valid for two further calls) -- the D2H copy of the actions and, when the caller feeds them back as
        ``u``, the H2D copy of the next step are then asynchronous instead of staged by the driver (~40 us each at 8192
        roll-outs).  Measurements are read from the caller's array as it is (pass a ``capi.pinned_empty`` array for the same
        effect)."""
        self.i2c = i2c
        self._pinned = [capi.pinned_empty((i2c.B, i2c.dims[1])) for _ in range(3)] if pinned_io else None
        self._n_calls = 0
        self.dim_u, self.dim_x = i2c.dims[1], i2c.dims[0]
        self.n_iter = int(n_iter)
        self.sig_u = np.asarray(sig_u, float)
        self.sig_zeta = sig_zeta
        # MpcPolicy.__init__ (policy/mpc.py:16-33): tau = 0; cell_init = deepcopy(cells[0]) BEFORE the targets are
        # assigned: the appended cells carry the initial action prior of cell 0 and the initial alpha (quirk A.6.6)
        i2c.tau = 0
        self._mu_u_init = i2c.mu_u_init[0, 0].copy()
        self._alpha_init = float(i2c.alpha0[0])
        self.z_traj = None if z_traj is None else np.asarray(z_traj, float)
        if self.z_traj is not None:
            self._set_targets(self.z_traj[: i2c.H])
        self._z_last = None if self.z_traj is None else self.z_traj[i2c.H - 1].copy()

    def _set_targets(self, z):
        import ctypes as C

        g = self.i2c
        assert not g.z_per_problem
        g.z = capi.f64(z, (g.H, g.dims[2]))
        # cell targets live in one small device array; re-upload through the problem setter's z path
        capi.check(g.lib.i2c_set_cell_targets(g._h, capi.ptr(g.z)))

    def set_control(self, feedforward):
        """policy/mpc.py:35-41 (tau = 0: cells stay independent; tau = H: feedback after the first _update_priors)."""
        self.i2c.tau = 0 if feedforward else self.i2c.H

    def filter(self, y, u):
        """PartiallyObservedMpcPolicy.filter (policy/mpc.py:125-145)."""
        self.i2c.ckf_step(y, u, self.sig_zeta)

    def optimize(self, n_iter, mu=None, covar=None):
        """policy/mpc.py:147-154: n_iter x (_forward_backward_msgs; _update_priors) from the current belief."""
        if mu is not None:
            self.i2c.set_initial_state(mu, covar)
        self.i2c.forward_backward(n_iter, update_priors=True)

    @property
    def belief(self):
        return self.i2c.get_initial_state()

    def _next_target(self, i):
        g = self.i2c
        if self.z_traj is None:
            return g.z_graph
        z_new = self.z_traj[i + g.H] if (i + g.H) < self.z_traj.shape[0] else self._z_last
        self._z_last = np.array(z_new, float)
        return z_new

    def __call__(self, i, y, u, fused=True):
        """policy/mpc.py:156-182 with deterministic=True: returns the first planned action [B, du].
        fused=True runs filter, sweeps, action read-out and horizon shift as ONE library call with a single
        synchronisation (i2c_mpc_step); fused=False issues the individual calls (same numbers)."""
        g = self.i2c
        z_new = capi.f64(self._next_target(i), (g.dims[2],))
        if not fused:
            if i > 0:
                self.filter(y, u)
            self.optimize(self.n_iter)
            ctrl, _ = g.first_action()
            g.shift_horizon(z_new, self._mu_u_init, self._alpha_init)
            return ctrl
        ctrl = np.empty((g.B, g.dims[1])) if self._pinned is None else self._pinned[self._n_calls % 3]
        self._n_calls += 1
        yy = uu = sz = None
        if i > 0:
            yy = capi.f64(np.broadcast_to(y, (g.B, g.env.dim_y)))
            uu = capi.f64(np.broadcast_to(u, (g.B, g.dims[1])))
            sz = capi.f64(self.sig_zeta, (g.env.dim_y, g.env.dim_y))
        capi.check(g.lib.i2c_set_tau(g._h, int(g.tau)))
        capi.check(g.lib.i2c_mpc_step(g._h, int(i > 0), capi.ptr(yy), capi.ptr(uu), capi.ptr(sz), self.n_iter,
                                      capi.ptr(z_new), capi.ptr(capi.f64(self._mu_u_init)), self._alpha_init,
                                      capi.ptr(ctrl)))
        return ctrl
