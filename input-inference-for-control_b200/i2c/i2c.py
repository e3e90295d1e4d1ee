"""``I2cGraph`` / ``I2cCell`` mirror (reference: i2c/i2c.py:51-1401) on the CUDA path.

``I2cGraph`` keeps the reference's constructor signature, method names and metric lists; every sweep is one call
into ``libi2c_b200.so`` (B = 1 problem, horizon H) through ``i2c_b200.BatchedI2c``.  ``I2cCell`` objects are lazy
views: reading ``cell.mu_x0_m`` fetches the per-cell record from the device (cached until the next sweep),
writing ``cell.z`` / ``cell.state_action_independence`` / ``cell.use_expert_controller`` is pushed to the device
before the next sweep.  Plotting methods of the reference (i2c.py:1403-1818) are out of scope.
"""
import copy
import logging
import os

import numpy as np

import i2c_b200
from i2c_b200 import capi
from i2c.exp_types import CubatureQuadrature, GaussHermiteQuadrature, Linearize
from i2c.inference.quadrature import QuadratureInference

PLOT_TIKZ = False  # i2c/i2c.py:18 (imported by scripts/lqr_compare.py:12, nonlinear_covariance_control.py:6)

# cell attribute -> (device field, slicer); x = state block, u = action block of a joint quantity
_VEC = {
    "mu_xu0_m": ("mu_xu0_m", None), "mu_xu1_m": ("mu_xu0_m", None), "mu_x0_m": ("mu_xu0_m", "x"), "mu_u0_m": ("mu_xu0_m", "u"),
    "mu_xu1_f": ("mu_xu1_f", None), "mu_x1_f": ("mu_xu1_f", "x"), "mu_u1_f": ("mu_xu1_f", "u"),
    "mu_x3_f": ("mu_x3_f", None), "mu_z0_f": ("mu_z0_f", None), "mu_z0_m": ("mu_z0_m", None), "mu_x3_m": ("mu_x3_m", None),
    "mu_xu0_pf": ("mu_xu0_pf", None), "mu_x0_pf": ("mu_xu0_pf", "x"), "mu_u0_pf": ("mu_xu0_pf", "u"),
    "mu_z0_pf": ("mu_z0_pf", None), "mu_x3_pf": ("mu_x3_pf", None), "k": ("k", None),
    "nu_x3_b": ("nu_x3_b", None), "nu_x0_b": ("nu_x0_b", None), "mu_xu0_f_prev": ("mu_xu0_f", None),
}
_MAT = {
    "sig_xu0_m": ("sig_xu0_m", None), "sig_xu1_m": ("sig_xu0_m", None), "sig_x0_m": ("sig_xu0_m", "xx"),
    "sig_u0_m": ("sig_xu0_m", "uu"), "sig_xu1_f": ("sig_xu1_f", None), "sig_x1_f": ("sig_xu1_f", "xx"),
    "sig_u1_f": ("sig_xu1_f", "uu"), "sig_x3_f": ("sig_x3_f", None), "sig_z0_f": ("sig_z0_f", None),
    "sig_z0_m": ("sig_z0_m", None), "sig_x3_m": ("sig_x3_m", None), "sig_xu0_pf": ("sig_xu0_pf", None),
    "sig_x0_pf": ("sig_xu0_pf", "xx"), "sig_u0_pf": ("sig_xu0_pf", "uu"), "sig_z0_pf": ("sig_z0_pf", None),
    "sig_x3_pf": ("sig_x3_pf", None), "K": ("K", None), "sigK": ("sigK", None), "J_dyn": ("J_dyn", None),
    "lambda_x3_b": ("lambda_x3_b", None), "lambda_x0_b": ("lambda_x0_b", None), "sig_xu0_f_prev": ("sig_xu0_f", None),
}


class I2cCell(object):
    """View of one timestep of the device-resident factor graph."""

    _LOCAL = ("_g", "index", "_z", "_indep", "_expert", "_terminal", "sys")

    def __init__(self, graph, i):
        object.__setattr__(self, "_g", graph)
        object.__setattr__(self, "index", i)
        object.__setattr__(self, "_z", np.array(graph.sys.zg, float).reshape(-1, 1))
        object.__setattr__(self, "_indep", True)
        object.__setattr__(self, "_expert", True)
        object.__setattr__(self, "_terminal", False)
        object.__setattr__(self, "sys", graph.sys)

    # ---- settable per-cell state -------------------------------------------------------------
    @property
    def z(self):
        return self._z

    @z.setter
    def z(self, v):
        object.__setattr__(self, "_z", np.array(v, float).reshape(-1, 1))
        self._g._dirty_targets = True

    @property
    def state_action_independence(self):
        self._g._pull_flags()
        return self._indep

    @state_action_independence.setter
    def state_action_independence(self, v):
        self._g._pull_flags()
        object.__setattr__(self, "_indep", bool(v))
        self._g._dirty_flags = True

    @property
    def use_expert_controller(self):
        return self._expert

    @use_expert_controller.setter
    def use_expert_controller(self, v):
        object.__setattr__(self, "_expert", bool(v))
        self._g._dirty_flags = True

    @property
    def terminal_cell(self):
        return self._terminal

    @terminal_cell.setter
    def terminal_cell(self, v):
        object.__setattr__(self, "_terminal", bool(v))
        self._g._dirty_flags = True

    def __setattr__(self, name, value):
        if name in self._LOCAL or isinstance(getattr(type(self), name, None), property):
            object.__setattr__(self, name, value)
        else:
            raise AttributeError(f"I2cCell.{name} is a read-only view of device state on the CUDA path")

    # ---- device-backed attributes ---------------------------------------------------------------
    def __getattr__(self, name):
        g = object.__getattribute__(self, "_g")
        t = g.cells.index(self)
        dx = g.sys.dim_x
        if name in ("mu_xu0_f", "sig_xu0_f"):
            # after _update_priors the reference's attribute holds the new prior (= posterior copy, i2c.py:1220)
            src = ("prior_mu" if name[0] == "m" else "prior_sig") if g._priors_fresh else name
            a = g._field(src)[t]
            return a.reshape(-1, 1) if name[0] == "m" else a
        if name in _VEC:
            f, sl = _VEC[name]
            a = g._field(f)[t]
            a = a[:dx] if sl == "x" else (a[dx:] if sl == "u" else a)
            return a.reshape(-1, 1)
        if name in _MAT:
            f, sl = _MAT[name]
            a = g._field(f)[t]
            return a[:dx, :dx] if sl == "xx" else (a[dx:, dx:] if sl == "uu" else a)
        if name == "Jx_dyn":
            return g._field("J_dyn")[t][:dx, :]
        if name in ("mu_x0_f", "sig_x0_f"):
            if t == 0:
                return np.array(g.sys.x0 if name[0] == "m" else g.sys.sig_x0, float)
            prev = g._field("mu_x3_f" if name[0] == "m" else "sig_x3_f")[t - 1]
            return prev.reshape(-1, 1) if name[0] == "m" else prev
        if name == "sig_x_lag_m":
            return g._field("J_dyn")[t][:dx, :] @ g._field("sig_x3_m")[t]
        if name in ("mu_z3_m", "sig_z3_m"):
            if t != len(g.cells) - 1 or g.sig_xi_terminal_base is None:
                return None
            a = g._field(name)[0]
            return a.reshape(-1, 1) if name[0] == "m" else a
        if name == "u_pol":
            return self.K @ self.mu_x0_m + self.k
        if name == "u_pol_K":
            return self.K @ self.mu_x0_m
        if name == "sig_xi":
            return g.alpha * g.sig_xi0
        if name == "lam_xi":
            return np.linalg.inv(g.alpha * g.sig_xi0)
        if name == "sig_xi_terminal":
            return g.sig_xi_terminal
        if name in ("sig_eta", "sig_eta_pf"):
            return np.array(g.sys.sig_eta, float)
        if name == "z_term":
            return np.array(g.sys.zg_term, float).reshape(-1, 1)
        if name == "temp":
            return g._g.temp
        if name in ("lambda_x3_f", "nu_x3_f"):
            lam = np.linalg.inv(g._field("sig_x3_f")[t])
            return lam if name[0] == "l" else lam @ g._field("mu_x3_f")[t].reshape(-1, 1)
        raise AttributeError(name)

    def expected_observation_covar(self):
        d = self.z - self.mu_z0_m
        return np.outer(d, d) + self.sig_z0_m

    def expected_propagated_observation_covar(self):
        d = self.z - self.mu_z0_pf
        return np.outer(d, d) + self.sig_z0_pf


class I2cGraph(object):
    """Gaussian i2c for a whole trajectory (constructor signature of i2c/i2c.py:735-750)."""

    def __init__(self, sys, horizon, Q, R, Qf, alpha, alpha_update_tol, mu_u, sig_u, mu_x_terminal, sig_x_terminal,
                 inference, res_dir=None):
        self.sys = sys
        self.H = horizon
        self.z = np.array(sys.zg, float).reshape(-1, 1)
        self.z_term = np.array(sys.zg_term, float).reshape(-1, 1)
        self.alpha_base = alpha
        self.alpha_update_tol = alpha_update_tol
        self.Q, self.R, self.inference = Q, R, inference
        self.QR = i2c_b200.batched._block_diag(Q, R)
        self.lam_xi0 = np.copy(self.QR)
        self.sig_xi0 = np.linalg.inv(self.QR)
        if Qf is not None:
            self.Qf = Qf
            self.sig_xi_terminal_base = np.linalg.inv(Qf)
        else:
            logging.info("Qf is None")
            self.Qf = np.zeros((sys.dim_x, sys.dim_x))
            self.sig_xi_terminal_base = None
        self.mu_x_terminal = None if mu_x_terminal is None else np.asarray(mu_x_terminal, float).reshape((sys.dim_x, 1))
        self.sig_x_terminal = sig_x_terminal
        assert np.linalg.det(self.sig_xi0) > 0.0
        if isinstance(inference, CubatureQuadrature):
            kind, quad = "cubature", (inference.alpha, inference.beta, inference.kappa)
            self.obs_inf = QuadratureInference(inference, sys.dim_xu)
        elif isinstance(inference, GaussHermiteQuadrature):  # i2c.py:116, 839
            kind, quad = "gauss_hermite", (int(inference.degree), 0.0, 0.0)
            self.obs_inf = QuadratureInference(inference, sys.dim_xu)
        elif isinstance(inference, Linearize):
            kind, quad = "linearize", (1.0, 0.0, 0.0)
            self.obs_inf = QuadratureInference(CubatureQuadrature(1, 0, 0), sys.dim_xu)
        else:
            raise ValueError("Unknown inference method")
        mu_u = np.asarray(mu_u, float).reshape(horizon, sys.dim_u)
        self._ctor = dict(env=sys._b200_env, horizon=horizon, Q=Q, R=R, Qf=Qf, alpha=alpha, tol=alpha_update_tol, mu_u=mu_u,
                          sig_u=np.asarray(sig_u, float), mu_x_terminal=mu_x_terminal, sig_x_terminal=sig_x_terminal,
                          kind=kind, quad=quad)
        self._build()
        self.cells = [I2cCell(self, i) for i in range(horizon)]
        self.cells[-1].terminal_cell = True
        self._cell_init_mu_u = mu_u[0].copy()
        self._propagate = False
        self.tau = horizon - 1
        self.res_dir = res_dir
        self.policy_valid = False
        self.alphas = [alpha]
        self.alphas_desired = [alpha]
        self.alphas_pf = [alpha]
        self.costs_m_all, self.costs_p_all, self.costs_pf_all = [], [], []
        self.reset_metrics(False)

    def _build(self):
        c, s = self._ctor, self.sys
        self._g = i2c_b200.BatchedI2c(
            c["env"], 1, c["horizon"], c["Q"], c["R"], c["Qf"], c["alpha"], c["tol"], c["mu_u"], c["sig_u"],
            c["mu_x_terminal"], c["sig_x_terminal"], x0=np.asarray(s.x0, float).reshape(1, -1), sig_x0=s.sig_x0,
            sig_eta=s.sig_eta, z=np.repeat(np.asarray(s.zg, float).reshape(1, -1), c["horizon"], 0),
            z_term=np.asarray(s.zg_term, float).reshape(-1), env_par=s._b200_env_par(), quadrature=c["quad"],
            enable_aux=True, inference=c["kind"], device=getattr(s, "device", 0))
        self._g.z_graph[:] = np.asarray(s.zg, float).reshape(-1)
        self._g.reset()
        self._cache = {}
        self._dirty_flags = self._dirty_targets = True
        self._flags_stale = False
        self._priors_fresh = False
        self._x0_pushed = None

    def close(self):
        self._g.close()

    # ---- host <-> device state -------------------------------------------------------------------
    def _field(self, name):
        if name not in self._cache:
            self._cache[name] = self._g.field(name)[0]
        return self._cache[name]

    def _pull_flags(self):
        if self._flags_stale:
            f = self._g.get_cell_flags()
            for c, v in zip(self.cells, f):
                object.__setattr__(c, "_indep", bool(v & capi.CELL_INDEPENDENT))
            self._flags_stale = False

    def _push(self):
        g = self._g
        if self._dirty_flags:
            self._pull_flags()
            own = g.get_cell_flags() & capi.CELL_OWN_ALPHA
            f = np.array([(capi.CELL_INDEPENDENT if c._indep else 0) | (capi.CELL_EXPERT if c._expert else 0)
                          | (capi.CELL_TERMINAL if c._terminal else 0) for c in self.cells], np.int32) | own
            g.set_cell_flags(f)
            self._dirty_flags = False
        if self._dirty_targets:
            z = np.stack([c._z.reshape(-1) for c in self.cells])
            capi.check(g.lib.i2c_set_cell_targets(g._h, capi.ptr(capi.f64(z, (self.H, self.sys.dim_z)))))
            self._dirty_targets = False
        x0 = (np.asarray(self.sys.x0, float).tobytes(), np.asarray(self.sys.sig_x0, float).tobytes())
        if x0 != self._x0_pushed:  # sys.x0 / sys.sig_x0 are re-read by every sweep (i2c.py:877-878)
            g.set_initial_state(np.asarray(self.sys.x0, float).reshape(1, -1), np.asarray(self.sys.sig_x0, float))
            self._x0_pushed = x0
        g.tau = self.tau
        g._propagate = self._propagate
        # extension (not in the reference): cells per chunk of the parallel-in-time sweep, None = sequential kernel
        g.time_parallel_chunk = getattr(self, "time_parallel_chunk", None)

    def _after(self, phases):
        self._cache = {}
        if phases & capi.PH_UPDATE_PRIORS:
            self._flags_stale = True
            self._priors_fresh = True
        elif phases & capi.PH_FORWARD:
            self._priors_fresh = False
        st, info = self._g.status()
        if st[0] != 0:
            it, cell = int(info[0]) >> 16, int(info[0]) & 0xFFFF
            name = capi.STATUS_NAMES[st[0]]
            # the device words are sticky: clear them so that a caller who catches the exception and repairs alpha / x0
            # does not see this failure again on the next sweep (the reference raises once, at the failing call)
            self._g.clear_status()
            msg = f"i2c failure {name} at EM iteration {it} of the call, cell {cell}"
            if name == "NAN_ALPHA":
                raise ValueError("Alpha is NaN")
            if name == "POLICY_DET":
                raise ValueError(msg)
            raise np.linalg.LinAlgError(msg + " (Matrix is not positive definite)")

    def _run(self, n, phases):
        self._push()
        self._g.run(n, phases, collect=False)
        self._after(phases)

    @property
    def alpha(self):
        return float(self._g.alpha[0])

    @alpha.setter
    def alpha(self, v):
        self._g.alpha = float(v)

    @property
    def sig_xi(self):
        return self.alpha * self.sig_xi0

    @property
    def sig_xi_terminal(self):
        return None if self.sig_xi_terminal_base is None else self.alpha * self.sig_xi_terminal_base

    # ---- sweeps (names of the reference) ---------------------------------------------------------------
    def _forward_msgs(self):
        self._run(1, capi.PH_FORWARD)

    def _backward_msgs(self):
        self._run(1, capi.PH_BACKWARD)

    def _forward_backward_msgs(self):
        self._run(1, capi.PH_FORWARD | capi.PH_BACKWARD)

    def _update_priors(self):
        self._run(1, capi.PH_UPDATE_PRIORS)

    def _backward_ricatti_msgs(self):
        self._run(1, capi.PH_RICCATI)
        self.policy_valid = True

    def propagate(self):
        self._run(1, capi.PH_PROPAGATE)

    def _metric(self, name, n=1):
        buf = np.empty((n, 1))
        capi.check(self._g.lib.i2c_get_metric(self._g._h, capi.METRICS[name], capi.ptr(buf), n))
        return [float(v) for v in buf[:, 0]]

    def _metrics(self, names, n=1):
        """Several metrics of the last ``n`` iterations with ONE synchronisation (i2c_get_metrics): {name: [n floats]}."""
        ids = np.array([capi.METRICS[m] for m in names], np.int32)
        buf = np.empty((len(names), n, 1))
        capi.check(self._g.lib.i2c_get_metrics(self._g._h, capi.ptr(ids), len(names), capi.ptr(buf), n))
        return {m: [float(v) for v in buf[i, :, 0]] for i, m in enumerate(names)}

    def calibrate_alpha(self, only_decrease=False):
        assert self._propagate
        before = self.alpha
        self._run(1, capi.PH_PROPAGATE | capi.PH_CALIBRATE | (capi.PH_ONLY_DECREASE if only_decrease else 0))
        logging.info(f"calibrating alpha from propagation {before}->{self.alpha}")
        self.alphas[-1] = self.alpha

    def learn_msgs(self, n_iter=1):
        """E step (forward, backward, optional propagate) + M step; ``n_iter`` > 1 fuses several EM iterations in
        one kernel launch (extension; the reference runs one per call)."""
        ph = capi.PH_LEARN | (capi.PH_PROPAGATE if self._propagate else 0)
        done = 0
        while done < n_iter:
            n = min(n_iter - done, self._g.max_iters)
            self._run(n, ph)
            self.em_iter += n
            names = ["cost_m", "cost_m_var", "alpha_desired", "alpha", "policy_entropy", "x_prior_entropy"]
            if self._propagate:
                names += ["cost_pf", "cost_pf_var", "cost_pf_min", "alpha_pf", "propagate_entropy"]
                if self.sig_x_terminal is not None and self.mu_x_terminal is not None:
                    names += ["kl_term"]
            m = self._metrics(names, n)  # one read-back for the whole iteration's bookkeeping (i2c.py:1004-1133)
            self.costs_m += m["cost_m"]
            self.costs_m_var += m["cost_m_var"]
            self.alphas_desired += m["alpha_desired"]
            self.alphas += m["alpha"]
            self.policy_entropy += m["policy_entropy"]
            xe = m["x_prior_entropy"]
            self.x_prior_entropy += xe
            self.x_prior_neg_entropy += [-v for v in xe]
            se = 0.5 * np.log(np.linalg.det(2 * np.pi * np.e * np.asarray(self.sys.sig_eta, float))) * self.H
            self.sig_eta_entropy += [float(se)] * n
            self.sig_eta_pf_entropy += [float(se)] * n
            if self._propagate:
                self.costs_pf += m["cost_pf"]
                self.costs_pf_var += m["cost_pf_var"]
                self.cost_pf_min += m["cost_pf_min"]
                self.alphas_pf += m["alpha_pf"]
                self.propagate_entropy += m["propagate_entropy"]
                if "kl_term" in m:
                    self.kl_terms += m["kl_term"]
            else:
                self.costs_pf += [-1.0] * n
            done += n

    # ---- alpha bookkeeping on the host (names of the reference; the sweeps update alpha on the device) ----------
    def get_z_covar(self):
        """sum_t E[(z_t - z)(z_t - z)^T] under the posterior (i2c.py:983-984)."""
        return sum(c.expected_observation_covar() for c in self.cells)

    def get_z_propagated_covar(self):
        return sum(c.expected_propagated_observation_covar() for c in self.cells)

    def get_z_terminal_covar(self):
        c = self.cells[-1]
        d = np.asarray(self.z_term, float).reshape(-1, 1) - np.asarray(c.mu_z3_m, float).reshape(-1, 1)
        return d @ d.T + c.sig_z3_m

    def calculate_alpha(self, z_covar, z_covar_term=None):
        tr, sf = np.trace(self.QR @ z_covar), float(self.sys.dim_z * self.H)
        if z_covar_term is not None:
            tr += np.trace(self.Qf @ z_covar_term)
            sf += float(self.sys.dim_z_term)
        return tr / sf

    def compute_update_alpha(self, update_alpha, *unused, **unused_kw):
        """i2c.py:921-946 on the current device messages (extra arguments of the call in policy/mpc.py:59 are accepted
        and ignored, the reference's own signature takes none)."""
        term = self.get_z_terminal_covar() if self.sig_xi_terminal_base is not None else None
        alpha_update = self.calculate_alpha(self.get_z_covar(), term)
        if self._propagate:
            self.alphas_pf.append(self.calculate_alpha(self.get_z_propagated_covar()))
        self.alphas_desired.append(alpha_update)
        if update_alpha:
            self.update_alpha(alpha_update)
        self.alphas.append(self.alpha)

    def update_alpha(self, alpha_update):
        if np.isnan(alpha_update):
            raise ValueError("Alpha is NaN")
        if self.alpha_update_tol >= 0.0:
            ratio, upper = alpha_update / self.alpha, 2.0 - self.alpha_update_tol
            if ratio < self.alpha_update_tol:
                alpha_update = self.alpha_update_tol * self.alpha
            if ratio > upper:
                alpha_update = upper * self.alpha
        else:
            alpha_update = self.alpha
        self._update_alpha(alpha_update)

    def _update_alpha(self, update):
        self.alpha = update  # i2c_set_alpha: sig_xi = alpha * QR^-1 is rebuilt by every cell on the device

    def _override_alpha(self, update):
        self.alphas[-1] = update
        self.alpha = update

    def update_xi(self, *unused):
        pass  # cells read alpha * sig_xi0 from the graph (device scalar); nothing to broadcast

    def update_models(self):
        pass

    @property
    def propagate_cost_improved(self):
        return self.costs_pf[-1] <= self.costs_pf[-2] if len(self.costs_pf) > 1 else True

    # ---- likelihood diagnostics (i2c.py:690-718, 1135-1170; never called by the reference's scripts) -------------
    def _calc_likelihood(self):
        """Expected complete-data log-likelihood terms of the current messages.  The per-cell dynamics moments
        (cell._calc_likelihood_quadrature, i2c.py:690-705) are one batched sigma-point launch over the cells; the
        remaining small traces are formed on the host from the device fields.  Quirks kept: the normalising terms use
        det(.) (not log det), the smoothed marginals are joined block-diagonally (concat_normals)."""
        s, H = self.sys, self.H
        dx = s.dim_x
        m, S = self._field("mu_xu0_m"), self._field("sig_xu0_m")
        Sbd = np.zeros_like(S)
        Sbd[:, :dx, :dx], Sbd[:, dx:, dx:] = S[:, :dx, :dx], S[:, dx:, dx:]
        kind = self._ctor["kind"]
        if kind == "linearize":
            if not hasattr(s, "A"):
                raise NotImplementedError("likelihood of the Linearize path needs the per-cell linearisation (linear envs only)")
            A, B, a = np.asarray(s.A, float), np.asarray(s.B, float), np.asarray(s.a, float).reshape(-1)
            mu_x = m[:, :dx] @ A.T + m[:, dx:] @ B.T + a
            sig_x = A @ S[:, :dx, :dx] @ A.T + B @ S[:, dx:, dx:] @ B.T
        else:
            kw = dict(gh_degree=int(self._ctor["quad"][0])) if kind == "gauss_hermite" else dict(quad=self._ctor["quad"])
            mu_x, sig_x, _, st = i2c_b200.quadrature(s._b200_env, "forward", m, Sbd, env_par=s._b200_env_par(),
                                                     device=getattr(s, "device", 0), **kw)
            if np.any(st != 0):
                raise np.linalg.LinAlgError("Matrix is not positive definite")
        sig_eta = np.asarray(s.sig_eta, float)
        Jx = self._field("J_dyn")[:, :dx, :]
        mu3, S3 = self._field("mu_x3_m"), self._field("sig_x3_m")
        lag = Jx @ S3
        M11 = np.einsum("ti,tj->tij", mu3, mu3) + S3
        M01 = np.einsum("ti,tj->tij", mu_x, mu3) + lag
        M00 = np.einsum("ti,tj->tij", mu_x, mu_x) + sig_x
        ll_xu = -0.5 * np.trace(np.sum(np.linalg.solve(sig_eta, M00 - M01 - np.transpose(M01, (0, 2, 1)) + M11), axis=0))
        lam_xi = np.linalg.inv(self.sig_xi)
        ll_z = -0.5 * np.trace(lam_xi @ self.get_z_covar())
        ll_const = -0.5 * H * (dx + s.dim_z) * np.log(2 * np.pi)
        ll_sig_xi = -0.5 * H * np.linalg.det(self.sig_xi)
        ll_sig_eta = -0.5 * H * np.linalg.det(sig_eta)
        sig_x0 = np.asarray(s.sig_x0, float)
        ll_sig_x0 = -0.5 * np.linalg.det(sig_x0)
        d0 = m[0, :dx] - np.asarray(s.x0, float).reshape(-1)
        ll_mu_x0 = -0.5 * np.trace(np.linalg.solve(sig_x0, np.outer(d0, d0) + S[0, :dx, :dx]))
        ll_state_action, ll_cost = ll_sig_eta + ll_xu, ll_sig_xi + ll_z
        return ll_const + ll_cost + ll_state_action + ll_sig_x0 + ll_mu_x0, ll_state_action, ll_cost, ll_xu

    def calc_likelihood(self):
        ll, _, ll_z, ll_xu = self._calc_likelihood()  # (the reference's tuple unpacking keeps the last ll_xu, i2c.py:1160)
        self.likelihoods.append(ll)
        self.likelihoods_xu.append(ll_xu)
        self.likelihoods_z.append(ll_z)
        self.risk.append(-2 * ll_xu / self.alpha)

    @staticmethod
    def list_minima(values, n_min, n_steps):
        """i2c.py:1172-1181: once more than n_min values exist, True iff the last n_steps steps all went down
        (None before that, as in the reference)."""
        if len(values) <= n_min:
            return None
        if not (len(values) > n_steps > 0):
            return False
        tail = np.asarray(values[-(n_steps + 1):], float)
        return bool(np.all(np.diff(tail) < 0))

    def likelihood_z_minima(self, n_min, n_steps):
        return self.list_minima(self.likelihoods_z, n_min, n_steps)

    def likelihood_xu_minima(self, n_min, n_steps):
        return self.list_minima(self.likelihoods_xu, n_min, n_steps)

    def get_prior_state_action_distribution(self):
        return self._g.field("mu_xu0_f")[0].copy(), self._g.field("sig_xu0_f")[0].copy()

    def get_propagated_state(self):
        dx = self.sys.dim_x
        return self._field("mu_xu0_pf")[:, :dx].copy(), self._field("sig_xu0_pf")[:, :dx, :dx].copy()

    def __getattr__(self, name):
        # plot_traj / plot_metrics / plot_alphas / plot_cost / ... (i2c.py:1403-1818): figures are outside the CUDA path;
        # the scripts' calls are accepted and logged so that they run unchanged
        if name.startswith("plot_"):
            def _no_plot(*a, **k):
                logging.info(f"I2cGraph.{name}: plotting is not part of the CUDA path (skipped)")
            return _no_plot
        raise AttributeError(name)

    # ---- getters ----------------------------------------------------------------------------------------
    def get_local_linear_policy(self):
        # (cached until the next sweep: scripts/i2c_run.py:89-98 calls both getters after every iteration)
        if "policy" not in self._cache:
            K, k, s = self._g.get_local_linear_policy()
            self._cache["policy"] = (K[0], k[0], s[0])
        K, k, s = self._cache["policy"]
        return K.copy(), k.copy(), s.copy()

    def get_local_expert_linear_policy(self):
        K, k, s = self.get_local_linear_policy()
        dx = self.sys.dim_x
        mu = self._field("mu_xu0_m")
        sig = self._field("sig_xu0_m")
        lam = np.linalg.inv(sig[:, :dx, :dx])
        return K, mu[:, dx:].copy(), s, mu[:, :dx].copy(), lam

    def get_marginal_input(self):
        return self._field("mu_xu0_m")[:, self.sys.dim_x:, None].copy()

    def get_marginal_state_action(self):
        return self._field("mu_xu0_m")[:, :, None].copy()

    def get_state_action_prior(self):
        return np.asarray([c.mu_xu0_f for c in self.cells])

    def get_marginal_state_action_distribution(self):
        return self._field("mu_xu0_m").copy(), self._field("sig_xu0_m").copy()

    def get_marginal_trajectory(self):
        return self._field("mu_xu0_m").copy()

    def get_marginal_observed_trajectory(self):
        return self._field("mu_z0_m").copy(), self.cells[-1].mu_z3_m

    def get_state_and_action(self):
        m = self._field("mu_xu0_m")
        dx = self.sys.dim_x
        return m[:, :dx, None].copy(), m[:, dx:, None].copy()

    def get_propagated_state_action(self):
        return self._field("mu_xu0_pf").copy(), self._field("sig_xu0_pf").copy()

    @staticmethod
    def indexed_confidence_bound(mu, sig, idx):
        std = 2.0 * np.sqrt(sig[:, idx, idx])
        return mu[:, idx] + std, mu[:, idx] - std

    def converged(self):
        if len(self.costs_m) > 2:
            return abs(self.costs_m[-1] - self.costs_m[-2]) / self.costs_m[-1] < 0.005
        return False

    def reset_metrics(self, extend=True):
        if extend:
            self.costs_m_all.extend(self.costs_m)
            self.costs_pf_all.extend(self.costs_pf)
        for name in ("costs_m", "costs_m_var", "costs_pf", "costs_pf_var", "cost_pf_min", "policy_entropy",
                     "sig_eta_entropy", "sig_eta_pf_entropy", "x_prior_entropy", "x_prior_neg_entropy",
                     "propagate_entropy", "kl_terms", "costs_p", "likelihoods", "likelihoods_xu", "likelihoods_z", "risk"):
            setattr(self, name, [])
        self.em_iter = 0

    def save_traj(self, res_dir):
        m = self._field("mu_xu0_m")
        dx = self.sys.dim_x
        np.save(os.path.join(res_dir, "xu_plan.npy"), m[:, :, None])
        np.save(os.path.join(res_dir, "x_plan.npy"), m[:, :dx, None])
        np.save(os.path.join(res_dir, "u_plan.npy"), m[:, dx:, None])
        np.save(os.path.join(res_dir, "z_plan.npy"), self._field("mu_z0_m")[:, :, None])

    # ---- deepcopy / pickle: snapshot of the device state -------------------------------------------------
    def __getstate__(self):
        # pending host-side cell edits (c.use_expert_controller = ..., c.z = ..., c.state_action_independence = ...) are
        # pushed lazily at the next sweep: push them now so that the copy runs with what its cell objects report
        if self._dirty_flags or self._dirty_targets:
            self._push()
        self._pull_flags()
        d = {k: v for k, v in self.__dict__.items() if k not in ("_g", "_cache", "cells")}
        d["_snapshot"] = self._g.snapshot()
        d["_cells"] = [(c.index, c._z, c._indep, c._expert, c._terminal) for c in self.cells]
        return d

    def __setstate__(self, d):
        snap, cells = d.pop("_snapshot"), d.pop("_cells")
        self.__dict__.update(d)
        self._build()
        self._g.restore(snap)
        self.cells = []
        for (idx, z, indep, expert, term) in cells:
            c = I2cCell(self, idx)
            for k, v in (("_z", z), ("_indep", indep), ("_expert", expert), ("_terminal", term)):
                object.__setattr__(c, k, v)
            self.cells.append(c)
        self._dirty_flags = self._dirty_targets = False
        self._x0_pushed = None

    def __deepcopy__(self, memo):
        new = object.__new__(type(self))
        state = self.__getstate__()
        state = {k: (v if k in ("sys", "_snapshot") else copy.deepcopy(v, memo)) for k, v in state.items()}
        new.__setstate__(state)
        return new

    def save(self, path, name):
        import pickle

        with open(os.path.join(path, f"i2c_{name}.pkl"), "wb") as f:
            pickle.dump(self, f)

    @classmethod
    def load(cls, path):
        import pickle

        with open(path, "rb") as f:
            return pickle.load(f)
