"""Batched Gaussian i2c on the GPU: thousands of independent trajectory-optimisation problems advanced by
one persistent CUDA kernel per call (csrc/i2c_kernels.cuh) through the C-ABI of include/i2c_b200.h.

``BatchedI2c`` is the batched counterpart of the reference's ``I2cGraph`` (i2c/i2c.py:732-1401): same
constructor arguments plus per-problem initial states, same method names for the sweeps, per-problem
status words instead of exceptions.  ``i2c.i2c.I2cGraph`` in this package (the drop-in mirror) is the
B = 1 view of this class.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from . import envs as _envs


def _block_diag(Q, R):
    if Q is None:
        return np.asarray(R, float)
    Q, R = np.asarray(Q, float), np.asarray(R, float)
    out = np.zeros((Q.shape[0] + R.shape[0],) * 2)
    out[: Q.shape[0], : Q.shape[0]] = Q
    out[Q.shape[0]:, Q.shape[0]:] = R
    return out


class BatchedI2c:
    def __init__(self, env, n_problems, horizon, Q, R, Qf, alpha, alpha_update_tol, mu_u, sig_u, mu_x_terminal=None,
                 sig_x_terminal=None, x0=None, sig_x0=None, sig_eta=None, z=None, z_term=None, env_par=None,
                 quadrature=(1.0, 0.0, 0.0), device=0, enable_aux=False, max_iters=256, dtemp=1.0, stream=None,
                 z_per_problem=False, torch_workspace=True, inference="cubature"):
        self.lib = capi.lib()
        self.env = _envs.make(env) if isinstance(env, str) else env
        self.env_id = capi.ENV_IDS[self.env.name]
        e = self.env
        self.B, self.H = int(n_problems), int(horizon)
        B, T = self.B, self.H
        dx, du, dz, dzt = e.dim_x, e.dim_u, e.dim_z, e.dim_z_term
        self.dims = (dx, du, dz, dzt)
        self.device = device
        self.enable_aux = bool(enable_aux)
        self.max_iters = int(max_iters)
        self.inference = inference
        inf_id = {"cubature": capi.INF_CUBATURE, "linearize": capi.INF_LINEARIZE, "gauss_hermite": capi.INF_GAUSS_HERMITE}[inference]
        if inference == "gauss_hermite":
            # GaussHermiteQuadrature(degree) (exp_types.py:52-68): `quadrature` is the degree (int or 1-tuple)
            deg = int(quadrature[0] if np.ndim(quadrature) else quadrature)
            quadrature = (float(deg), 0.0, 0.0)
        cfg = capi.Config(capi.ABI_VERSION, self.env_id, inf_id, B, T, self.max_iters, device, int(bool(z_per_problem)),
                          int(self.enable_aux), *map(float, quadrature))
        self._cfg = cfg
        self._ws = None
        ws_ptr, ws_bytes, stream_ptr = None, 0, None
        if torch_workspace:
            import torch  # plumbing only: device memory + stream

            if not torch.cuda.is_available():
                raise capi.I2cError("no CUDA device available: the i2c hot path has no CPU fallback")
            nbytes = C.c_size_t()
            capi.check(self.lib.i2c_workspace_bytes(C.byref(cfg), C.byref(nbytes)))
            with torch.cuda.device(device):
                self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=f"cuda:{device}")
                s = stream if stream is not None else torch.cuda.current_stream(device)
            ws_ptr, ws_bytes, stream_ptr = self._ws.data_ptr(), nbytes.value, s.cuda_stream
        self._h = C.c_void_p()
        capi.check(self.lib.i2c_create(C.byref(cfg), C.c_void_p(ws_ptr), ws_bytes, C.c_void_p(stream_ptr), C.byref(self._h)))
        self.workspace_bytes = ws_bytes
        # ---- problem definition
        self.QR = _block_diag(Q, R)
        assert self.QR.shape == (dz, dz), (self.QR.shape, dz)
        self.Qf = None if Qf is None else capi.f64(Qf, (dzt, dzt))
        self.alpha_update_tol = float(alpha_update_tol)
        x0 = e.x0 if x0 is None else x0
        self.x0 = capi.f64(np.broadcast_to(np.asarray(x0, float).reshape((-1, dx)) if np.ndim(x0) <= 2 else x0, (B, dx)).copy())
        sig_x0 = e.sig_x0 if sig_x0 is None else sig_x0
        self.sig_x0 = capi.f64(np.broadcast_to(sig_x0, (B, dx, dx)).copy())
        self.sig_eta = capi.f64(e.sig_eta if sig_eta is None else sig_eta, (dx, dx))
        mu_u = np.asarray(mu_u, float)
        if mu_u.ndim == 2:
            mu_u = np.broadcast_to(mu_u[None], (B, T, du))
        self.mu_u_init = capi.f64(mu_u, (B, T, du))
        self.sig_u = capi.f64(sig_u, (du, du))
        self.z_graph = capi.f64(np.asarray(e.zg, float).reshape(-1), (dz,))
        if z is None:
            z = np.broadcast_to(self.z_graph, (T, dz))
        self.z_per_problem = bool(z_per_problem)
        self.z = capi.f64(z, (B, T, dz) if z_per_problem else (T, dz))
        zt = np.asarray(e.zg_term if z_term is None else z_term, float)
        self.z_term = capi.f64(np.broadcast_to(zt.reshape((-1, dzt)), (B, dzt)).copy()) if z_per_problem else capi.f64(zt.reshape(-1), (dzt,))
        self.alpha0 = capi.f64(np.broadcast_to(np.asarray(alpha, float), (B,)).copy())
        self.mu_x_terminal = None if mu_x_terminal is None else capi.f64(np.asarray(mu_x_terminal, float).reshape(-1), (dx,))
        self.sig_x_terminal = None if sig_x_terminal is None else capi.f64(sig_x_terminal, (dx, dx))
        n_par = capi.env_dims(self.env_id)[4]
        if n_par:
            if env_par is None:
                env_par = _envs.linear_params(e.A, e.B, e.a)
            env_par = capi.f64(np.broadcast_to(env_par, (B, n_par)).copy())
        self.env_par = env_par
        self.dtemp = float(dtemp)
        self._propagate = False
        self.tau = T - 1
        self.time_parallel_chunk = None  # cells per chunk of the parallel-in-time sweep (None = sequential kernel)
        self.reset()

    # ------------------------------------------------------------------ lifecycle
    def reset(self):
        """(Re)initialise every cell to its constructor state (I2cCell.__init__, i2c.py:54-148)."""
        p = capi.ptr
        capi.check(self.lib.i2c_set_problem(
            self._h, p(self.x0), p(self.sig_x0), p(self.sig_eta), p(self.mu_u_init), p(self.sig_u), p(capi.f64(self.QR)),
            p(self.Qf), p(self.z), p(self.z_graph), p(self.z_term), p(self.alpha0), self.alpha_update_tol,
            p(self.mu_x_terminal), p(self.sig_x_terminal), self.dtemp, p(self.env_par)))
        self.tau = self.H - 1
        self.metrics = {k: [] for k in capi.METRICS}
        self.alphas = [self.alpha0.copy()]
        self.em_iter = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.i2c_destroy(self._h)
            self._h = C.c_void_p()
        self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ sweeps
    def run(self, n_iter, phases, collect=True):
        """Launch the persistent kernel for `n_iter` iterations of the selected phases (asynchronous)."""
        if self.enable_aux:
            phases |= capi.PH_STORE_AUX
        done = 0
        while done < n_iter:
            n = min(self.max_iters, n_iter - done)
            capi.check(self.lib.i2c_set_tau(self._h, int(self.tau)))
            if self.time_parallel_chunk:
                # parallel-in-time sweep (csrc/i2c_scan.cuh): Linearize inference on linear systems only, refused otherwise
                capi.check(self.lib.i2c_run_scan(self._h, n, phases, int(self.time_parallel_chunk)))
            else:
                capi.check(self.lib.i2c_run(self._h, n, phases))
            if collect and (phases & (capi.PH_MSTEP | capi.PH_CALIBRATE | capi.PH_PROPAGATE)):
                self._collect(n, phases)
            done += n
        return self

    def _collect(self, n, phases):
        buf = np.empty((n, self.B))
        names = []
        if phases & capi.PH_MSTEP:
            names += ["alpha", "alpha_desired", "cost_m", "cost_m_var", "cost_pf", "policy_entropy", "x_prior_entropy"]
            if phases & capi.PH_PROPAGATE:
                names += ["alpha_pf", "cost_pf_var", "cost_pf_min", "propagate_entropy"]
        elif phases & capi.PH_CALIBRATE:
            names += ["alpha"]
        if (phases & capi.PH_PROPAGATE) and self.sig_x_terminal is not None and (phases & capi.PH_MSTEP):
            names += ["kl_term"]
        for name in names:
            capi.check(self.lib.i2c_get_metric(self._h, capi.METRICS[name], capi.ptr(buf), n))
            self.metrics[name].extend(buf.copy())
        if phases & capi.PH_MSTEP:
            self.alphas.extend(self.metrics["alpha"][-n:])
        elif phases & capi.PH_CALIBRATE:
            self.alphas[-1] = self.metrics["alpha"][-1]  # _override_alpha (i2c.py:970-974)

    def learn(self, n_iter=1, collect=True):
        """n_iter x I2cGraph.learn_msgs (i2c.py:1238-1245)."""
        ph = capi.PH_LEARN | (capi.PH_PROPAGATE if self._propagate else 0)
        self.em_iter += n_iter
        return self.run(n_iter, ph, collect)

    learn_msgs = learn

    def forward_backward(self, n_iter=1, update_priors=False):
        """_forward_backward_msgs (i2c.py:1231-1236), optionally followed by _update_priors (MPC optimise loop)."""
        return self.run(n_iter, capi.PH_FORWARD | capi.PH_BACKWARD | (capi.PH_UPDATE_PRIORS if update_priors else 0), False)

    def forward(self):
        return self.run(1, capi.PH_FORWARD, False)

    def backward(self):
        return self.run(1, capi.PH_BACKWARD, False)

    def update_priors(self):
        return self.run(1, capi.PH_UPDATE_PRIORS, False)

    def propagate(self):
        return self.run(1, capi.PH_PROPAGATE, False)

    def backward_ricatti(self):
        """I2cGraph._backward_ricatti_msgs (i2c.py:888-893): Linearize inference on linear envs, enable_aux=True."""
        return self.run(1, capi.PH_RICCATI, False)

    def calibrate_alpha(self, only_decrease=False):
        """I2cGraph.calibrate_alpha (i2c.py:895-911)."""
        return self.run(1, capi.PH_PROPAGATE | capi.PH_CALIBRATE | (capi.PH_ONLY_DECREASE if only_decrease else 0))

    def synchronize(self):
        capi.check(self.lib.i2c_synchronize(self._h))

    # ------------------------------------------------------------------ state access
    @property
    def alpha(self):
        out = np.empty(self.B)
        capi.check(self.lib.i2c_get_alpha(self._h, capi.ptr(out)))
        return out

    @alpha.setter
    def alpha(self, value):
        a = capi.f64(np.broadcast_to(np.asarray(value, float), (self.B,)).copy())
        capi.check(self.lib.i2c_set_alpha(self._h, capi.ptr(a)))

    def status(self):
        st, info = np.zeros(self.B, np.int32), np.zeros(self.B, np.int32)
        capi.check(self.lib.i2c_get_status(self._h, capi.ptr(st), capi.ptr(info)))
        return st, info

    def clear_status(self):
        """Reset the (sticky) per-problem status / info words, e.g. after a failure has been handled."""
        capi.check(self.lib.i2c_clear_status(self._h))

    def field(self, name, t0=0, t1=None):
        """Per-cell attribute for cells [t0, t1): [B, t1-t0, rows(, cols)] (vectors squeezed)."""
        fid = capi.FIELDS[name]
        r, c = C.c_int32(), C.c_int32()
        capi.check(self.lib.i2c_field_shape(self._h, fid, C.byref(r), C.byref(c)))
        t1 = self.H if t1 is None else t1
        if name in ("mu_z3_m", "sig_z3_m"):
            t0, t1 = 0, 1
        out = np.empty((self.B, t1 - t0, r.value, c.value))
        capi.check(self.lib.i2c_get_field(self._h, fid, t0, t1, capi.ptr(out)))
        is_vec = c.value == 1 and not name.startswith(("sig", "K", "J", "lambda"))
        if name == "K" or name == "prior_K" or name == "J_dyn":
            is_vec = False
        return out[..., 0] if is_vec else out

    def set_field(self, name, value, t0=0, t1=None):
        fid = capi.FIELDS[name]
        r, c = C.c_int32(), C.c_int32()
        capi.check(self.lib.i2c_field_shape(self._h, fid, C.byref(r), C.byref(c)))
        t1 = self.H if t1 is None else t1
        v = capi.f64(np.asarray(value, float).reshape(self.B, t1 - t0, r.value, c.value))
        capi.check(self.lib.i2c_set_field(self._h, fid, t0, t1, capi.ptr(v)))

    def get_local_linear_policy(self):
        """I2cGraph.get_local_linear_policy (i2c.py:1253-1264), batched: K[B,H,du,dx], k[B,H,du], sigK[B,H,du,du]."""
        dx, du = self.dims[0], self.dims[1]
        K = np.empty((self.B, self.H, du, dx))
        k = np.empty((self.B, self.H, du))
        s = np.empty((self.B, self.H, du, du))
        capi.check(self.lib.i2c_get_policy(self._h, capi.ptr(K), capi.ptr(k), capi.ptr(s)))
        return K, k, s

    def get_local_linear_policy_async(self, K, k, sigK):
        """Start the device->host copy of the controllers into the given (pinned) arrays on the copy stream; it
        overlaps later sweeps.  Call wait_copies() before reading the arrays."""
        capi.check(self.lib.i2c_get_policy_async(self._h, capi.ptr(K), capi.ptr(k), capi.ptr(sigK)))

    def wait_copies(self):
        capi.check(self.lib.i2c_copy_wait(self._h))

    def last_metrics_async(self, names, out, slot):
        """Queue the read-back of the given metrics of the LAST iteration of the most recent sweep into the page-locked array
        ``out`` [len(names), B] (capi.pinned_empty) through staging slot 0 / 1 and return at once; ``metrics_wait(slot)``
        blocks until the numbers have landed.  Lets a logging loop (scripts/i2c_run.py:84-88) queue step i+1 before it
        collects step i."""
        ids = np.array([capi.METRICS[n] for n in names], np.int32)
        assert out.shape == (len(names), self.B) and out.dtype == np.float64 and out.flags.c_contiguous
        capi.check(self.lib.i2c_get_last_metrics_async(self._h, capi.ptr(ids), len(names), capi.ptr(out), int(slot)))

    def metrics_wait(self, slot):
        capi.check(self.lib.i2c_metrics_wait(self._h, int(slot)))

    def get_cell_flags(self):
        f = np.zeros(self.H, np.int32)
        capi.check(self.lib.i2c_get_cell_flags(self._h, capi.ptr(f)))
        return f

    def set_cell_flags(self, flags):
        f = np.ascontiguousarray(flags, np.int32)
        assert f.shape == (self.H,)
        capi.check(self.lib.i2c_set_cell_flags(self._h, capi.ptr(f)))

    def set_cell_flag(self, flag, value, cells=None):
        f = self.get_cell_flags()
        idx = slice(None) if cells is None else cells
        f[idx] = (f[idx] | flag) if value else (f[idx] & ~flag)
        self.set_cell_flags(f)

    def set_initial_state(self, x0, sig_x0):
        x0 = capi.f64(np.broadcast_to(x0, (self.B, self.dims[0])).copy())
        s0 = capi.f64(np.broadcast_to(sig_x0, (self.B, self.dims[0], self.dims[0])).copy())
        capi.check(self.lib.i2c_set_initial_state(self._h, capi.ptr(x0), capi.ptr(s0)))

    def get_initial_state(self):
        dx = self.dims[0]
        x0, s0 = np.empty((self.B, dx)), np.empty((self.B, dx, dx))
        capi.check(self.lib.i2c_get_initial_state(self._h, capi.ptr(x0), capi.ptr(s0)))
        return x0, s0

    @property
    def temp(self):
        t = C.c_double()
        capi.check(self.lib.i2c_get_temp(self._h, C.byref(t)))
        return t.value

    @temp.setter
    def temp(self, v):
        capi.check(self.lib.i2c_set_temp(self._h, float(v)))

    # ------------------------------------------------------------------ MPC support (policy/mpc.py)
    def shift_horizon(self, z_new, mu_u_init, alpha_init):
        z_new = capi.f64(z_new)
        mu = capi.f64(np.asarray(mu_u_init, float).reshape(-1), (self.dims[1],))
        capi.check(self.lib.i2c_shift_horizon(self._h, capi.ptr(z_new), capi.ptr(mu), float(alpha_init)))

    def ckf_step(self, y, u, sig_zeta):
        dy = self.env.dim_y
        y = capi.f64(np.broadcast_to(y, (self.B, dy)).copy())
        u = capi.f64(np.broadcast_to(u, (self.B, self.dims[1])).copy())
        capi.check(self.lib.i2c_ckf_step(self._h, capi.ptr(y), capi.ptr(u), capi.ptr(capi.f64(sig_zeta, (dy, dy)))))

    def first_action(self):
        du = self.dims[1]
        mu, sig = np.empty((self.B, du)), np.empty((self.B, du, du))
        capi.check(self.lib.i2c_get_first_action(self._h, capi.ptr(mu), capi.ptr(sig)))
        return mu, sig

    # ------------------------------------------------------------------ snapshot (deepcopy / pickle support)
    def snapshot(self):
        n = C.c_size_t()
        capi.check(self.lib.i2c_snapshot_bytes(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        capi.check(self.lib.i2c_snapshot(self._h, capi.ptr(buf), n.value))
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        capi.check(self.lib.i2c_restore(self._h, capi.ptr(buf), buf.size))

    # ------------------------------------------------------------------ introspection
    def kernel_launches(self):
        n = C.c_int64()
        capi.check(self.lib.i2c_kernel_launches(self._h, C.byref(n)))
        return n.value

    def last_run_ms(self):
        ms = C.c_float()
        capi.check(self.lib.i2c_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def policy_device_tensors(self):
        """Controllers as torch CUDA tensors in canonical layout (input of the NCCL gather)."""
        import torch

        dx, du = self.dims[0], self.dims[1]
        dev = f"cuda:{self.device}"
        K = torch.empty((self.B, self.H, du, dx), dtype=torch.float64, device=dev)
        k = torch.empty((self.B, self.H, du), dtype=torch.float64, device=dev)
        s = torch.empty((self.B, self.H, du, du), dtype=torch.float64, device=dev)
        capi.check(self.lib.i2c_get_policy_dev(self._h, C.c_void_p(K.data_ptr()), C.c_void_p(k.data_ptr()),
                                               C.c_void_p(s.data_ptr())))
        return K, k, s


def gauss_hermite(degree):
    """The library's 1-D Gauss-Hermite rule: nodes, weights / sqrt(pi) (exp_types.py:57, 66)."""
    x, w = np.empty(degree), np.empty(degree)
    capi.check(capi.lib().i2c_gauss_hermite(int(degree), capi.ptr(x), capi.ptr(w)))
    return x, w


def quadrature(env, fn, m, S, quad=(1.0, 0.0, 0.0), env_par=None, device=0, gh_degree=None):
    """Stand-alone sigma-point transform on the GPU (QuadratureInference.forward / forward_gaussian,
    inference/quadrature.py:27-58) for a registered env map.  fn in {"observe", "observe_terminal", "forward",
    "measure"}.  m [B,d], S [B,d,d] -> m_y [B,dy], S_y [B,dy,dy], S_xy [B,d,dy], status [B].
    gh_degree: GaussHermiteQuadrature(degree) grid instead of the CubatureQuadrature(*quad) points."""
    L = capi.lib()
    env_id = capi.ENV_IDS[env]
    fn_id = {"observe": 0, "observe_terminal": 1, "forward": 2, "measure": 3}[fn]
    dx, du, dz, dzt, n_par, dy_meas = capi.env_dims(env_id)
    n = dx + du
    D = n if fn_id in (0, 2) else dx
    DY = [dz, dzt, dx, dy_meas][fn_id]
    m = capi.f64(np.atleast_2d(m))
    B = m.shape[0]
    m = capi.f64(m, (B, D))
    S = capi.f64(np.broadcast_to(S, (B, D, D)).copy())
    if n_par:
        if env_par is None:
            e = _envs.make(env)
            env_par = _envs.linear_params(e.A, e.B, e.a)
        env_par = capi.f64(np.broadcast_to(env_par, (B, n_par)).copy())
    my, Sy, Sxy = np.empty((B, DY)), np.empty((B, DY, DY)), np.empty((B, D, DY))
    st = np.zeros(B, np.int32)
    if gh_degree:
        capi.check(L.i2c_quadrature_gh(env_id, fn_id, B, capi.ptr(m), capi.ptr(S), int(gh_degree), capi.ptr(env_par),
                                       capi.ptr(my), capi.ptr(Sy), capi.ptr(Sxy), capi.ptr(st), device))
    else:
        capi.check(L.i2c_quadrature(env_id, fn_id, B, capi.ptr(m), capi.ptr(S), *map(float, quad), capi.ptr(env_par),
                                    capi.ptr(my), capi.ptr(Sy), capi.ptr(Sxy), capi.ptr(st), device))
    return my, Sy, Sxy, st


def rollout(env, x_init, K, k, sig_k=None, expert=None, soft_expert=True, eta=None, eps_u=None, sig_eta=None, seed=0,
            env_par=None, device=0, return_final=False):
    """Batched stochastic closed-loop evaluation of time-indexed linear-Gaussian controllers on the GPU
    (BaseSim.run / batch_eval, i2c/env.py:40-103, for all problems and roll-outs in one launch).

    x_init [B,R,dx]; K [B,T,du,dx], k [B,T,du]; sig_k [B,T,du,du] -> stochastic actions; expert = (mu [B,T,dx],
    lam [B,T,dx,dx]) -> ExpertTimeIndexedLinearGaussianPolicy; eta [B,R,T,dx] disturbances to add (else drawn on the
    device from N(0, sig_eta), default the env's); eps_u [B,R,T,du] standard normals for the action noise.
    Returns xu [B,R,T,dx+du], z [B,R,T,dz], z_term [B,R,dzt]."""
    L = capi.lib()
    env_id = capi.ENV_IDS[env]
    dx, du, dz, dzt, n_par, _ = capi.env_dims(env_id)
    e = _envs.make(env)
    x_init = capi.f64(x_init)
    B, R = x_init.shape[:2]
    K = capi.f64(K)
    T = K.shape[1]
    K = capi.f64(K, (B, T, du, dx))
    k = capi.f64(k, (B, T, du))
    x_init = capi.f64(x_init, (B, R, dx))
    sig_k = None if sig_k is None else capi.f64(sig_k, (B, T, du, du))
    mu = lam = None
    if expert is not None:
        mu, lam = capi.f64(expert[0], (B, T, dx)), capi.f64(expert[1], (B, T, dx, dx))
    eta = None if eta is None else capi.f64(eta, (B, R, T, dx))
    eps_u = None if eps_u is None else capi.f64(eps_u, (B, R, T, du))
    sig_eta = capi.f64(e.sig_eta if sig_eta is None else sig_eta, (dx, dx))
    if n_par:
        if env_par is None:
            env_par = _envs.linear_params(e.A, e.B, e.a)
        env_par = capi.f64(np.broadcast_to(env_par, (B, n_par)).copy())
    xu, z, zt = np.empty((B, R, T, dx + du)), np.empty((B, R, T, dz)), np.empty((B, R, dzt))
    xf = np.empty((B, R, dx))
    p = capi.ptr
    capi.check(L.i2c_rollout(env_id, B, R, T, p(x_init), p(K), p(k), p(sig_k), p(mu), p(lam), int(bool(soft_expert)), p(eta),
                             p(eps_u), p(sig_eta), int(seed), p(env_par), p(xu), p(z), p(zt), p(xf), device))
    return (xu, z, zt, xf) if return_final else (xu, z, zt)
