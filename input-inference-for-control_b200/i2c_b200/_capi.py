"""ctypes binding of include/i2c_b200.h (the C-ABI of the CUDA library).

The product path has NO CPU fallback: if the shared library is missing (not built) or no CUDA device is
visible, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("I2C_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libi2c_b200.so")  # env: A/B builds
ABI_VERSION = 1

# enum i2c_env
ENV_IDS = {
    "LinearKnown": 0,
    "LinearKnownMinimumEnergy": 1,
    "PendulumKnown": 2,
    "PendulumKnownActReg": 3,
    "CartpoleKnown": 4,
    "DoubleCartpoleKnown": 5,
    "Quadrotor": 6,
}
# enum i2c_phase
PH_FORWARD, PH_BACKWARD, PH_PROPAGATE, PH_MSTEP, PH_UPDATE_PRIORS = 1, 2, 4, 8, 16
PH_CALIBRATE, PH_ONLY_DECREASE, PH_STORE_AUX, PH_RICCATI = 32, 64, 128, 256
INF_CUBATURE, INF_LINEARIZE, INF_GAUSS_HERMITE = 0, 1, 2
PH_LEARN = PH_FORWARD | PH_BACKWARD | PH_MSTEP | PH_UPDATE_PRIORS
# enum i2c_cell_flag
CELL_INDEPENDENT, CELL_TERMINAL, CELL_EXPERT, CELL_OWN_ALPHA = 1, 2, 4, 8
# enum i2c_field
FIELDS = {name: i for i, name in enumerate([
    "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK", "prior_mu", "prior_sig", "prior_K",
    "mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn",
    "mu_xu0_f", "sig_xu0_f", "mu_z0_f", "sig_z0_f", "mu_z0_m", "sig_z0_m", "mu_x3_m", "sig_x3_m",
    "mu_xu0_pf", "sig_xu0_pf", "mu_z0_pf", "sig_z0_pf", "mu_x3_pf", "sig_x3_pf", "mu_z3_m", "sig_z3_m",
    "lambda_x3_b", "nu_x3_b", "lambda_x0_b", "nu_x0_b"])}
# enum i2c_metric
METRICS = {name: i for i, name in enumerate([
    "alpha", "alpha_desired", "alpha_pf", "cost_m", "cost_m_var", "cost_pf", "cost_pf_var", "cost_pf_min",
    "policy_entropy", "x_prior_entropy", "propagate_entropy", "kl_term"])}
STATUS_NAMES = ["OK", "CHOL_PRIOR", "CHOL_OBS", "CHOL_FILTERED", "CHOL_X3", "CHOL_TERMINAL", "CHOL_POSTERIOR", "MVN",
                "NAN_ALPHA", "POLICY_DET", "CHOL_PROPAGATE", "COV_CONTROL", "CKF"]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("env", C.c_int32), ("inference", C.c_int32), ("n_problems", C.c_int32),
                ("horizon", C.c_int32), ("max_iters", C.c_int32), ("device", C.c_int32), ("z_per_problem", C.c_int32),
                ("enable_aux", C.c_int32), ("quad_alpha", C.c_double), ("quad_beta", C.c_double),
                ("quad_kappa", C.c_double)]


class I2cError(RuntimeError):
    pass


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

EXPORTS = [
    "i2c_env_dims", "i2c_workspace_bytes", "i2c_create", "i2c_destroy", "i2c_set_problem", "i2c_set_initial_state", "i2c_set_initial_state_async",
    "i2c_set_initial_state_dev", "i2c_set_cell_flags", "i2c_get_cell_flags", "i2c_set_cell_index", "i2c_set_tau",
    "i2c_set_cell_targets",
    "i2c_set_alpha", "i2c_get_alpha", "i2c_set_temp", "i2c_get_temp", "i2c_run", "i2c_run_scan", "i2c_synchronize", "i2c_get_metric", "i2c_get_metrics",
    "i2c_get_status", "i2c_clear_status", "i2c_get_field", "i2c_set_field", "i2c_field_shape", "i2c_get_policy", "i2c_get_policy_async", "i2c_copy_wait", "i2c_get_policy_dev",
    "i2c_shift_horizon", "i2c_ckf_step", "i2c_mpc_step", "i2c_get_initial_state", "i2c_get_first_action", "i2c_quadrature", "i2c_quadrature_gh", "i2c_gauss_hermite",
    "i2c_rollout", "i2c_snapshot_bytes", "i2c_snapshot", "i2c_restore", "i2c_kernel_launches", "i2c_last_run_ms", "i2c_dfma_peak", "i2c_fastmath_probe",
    "i2c_host_alloc", "i2c_host_free", "i2c_get_last_metrics_async", "i2c_metrics_wait",
    "i2c_last_error",
    "i2c_build_info",
]


def lib():
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise I2cError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first. "
                       "There is no CPU fallback for the i2c hot path.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise I2cError(f"{LIB_PATH} does not export {name}")
    L.i2c_last_error.restype = C.c_char_p
    L.i2c_build_info.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("i2c_last_error", "i2c_build_info"):
            getattr(L, name).restype = C.c_int
    L.i2c_create.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
    L.i2c_workspace_bytes.argtypes = [C.POINTER(Config), C.POINTER(C.c_size_t)]
    L.i2c_destroy.argtypes = [C.c_void_p]
    L.i2c_set_problem.argtypes = [C.c_void_p] + [C.c_void_p] * 11 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                                                    C.c_void_p]
    L.i2c_set_initial_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_set_initial_state_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_set_initial_state_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_get_initial_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_set_cell_flags.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_get_cell_flags.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_set_cell_index.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_set_tau.argtypes = [C.c_void_p, C.c_int32]
    L.i2c_set_cell_targets.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_set_alpha.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_get_alpha.argtypes = [C.c_void_p, C.c_void_p]
    L.i2c_set_temp.argtypes = [C.c_void_p, C.c_double]
    L.i2c_get_temp.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.i2c_run.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.i2c_synchronize.argtypes = [C.c_void_p]
    L.i2c_get_metric.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    L.i2c_get_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    L.i2c_get_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_clear_status.argtypes = [C.c_void_p]
    L.i2c_get_field.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.i2c_set_field.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.i2c_field_shape.argtypes = [C.c_void_p, C.c_int32, _ip, _ip]
    L.i2c_get_policy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_get_policy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_copy_wait.argtypes = [C.c_void_p]
    L.i2c_get_policy_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_run_scan.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    L.i2c_shift_horizon.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.i2c_ckf_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_mpc_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_double, C.c_void_p]
    L.i2c_get_first_action.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.i2c_quadrature_gh.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.i2c_gauss_hermite.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
    L.i2c_quadrature.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.i2c_rollout.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 6 + [C.c_int32] + [C.c_void_p] * 3 + [C.c_uint64] + \
        [C.c_void_p] * 5 + [C.c_int32]
    L.i2c_snapshot_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.i2c_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.i2c_restore.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.i2c_kernel_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.i2c_last_run_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.i2c_dfma_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.i2c_fastmath_probe.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp]
    L.i2c_env_dims.argtypes = [C.c_int32, _ip, _ip, _ip, _ip, _ip, _ip]
    L.i2c_get_last_metrics_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    L.i2c_metrics_wait.argtypes = [C.c_void_p, C.c_int32]
    L.i2c_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.i2c_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise I2cError(f"i2c_b200 error {rc}: {lib().i2c_last_error().decode()}")


def f64(a, shape=None):
    """C-contiguous fp64 copy/view with an optional shape check."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype=np.float64):
    """NumPy array in page-locked host memory (i2c_host_alloc): copies from / to it are asynchronous.  The block is
    released when the last array viewing it is garbage-collected."""
    import weakref

    count = int(np.prod(shape))
    nbytes = max(count * np.dtype(dtype).itemsize, 1)
    p = C.c_void_p()
    check(lib().i2c_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)  # every array derived from it keeps `buf` alive
    weakref.finalize(buf, lib().i2c_host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def dfma_peak(device=0):
    v = C.c_double()
    check(lib().i2c_dfma_peak(device, C.byref(v)))
    return v.value


FASTMATH_FN = {"rsqrt": 0, "rcp": 1, "exp_neg": 2, "exp_neg_lat": 3, "sincos": 4, "seq_sincos": 5, "logacc": 6}


def fastmath_probe(fn, x, device=0):
    """Evaluate one device math primitive of csrc/fastmath.cuh on the array x; returns (y0, y1)."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    y0, y1 = np.empty_like(x), np.empty_like(x)
    check(lib().i2c_fastmath_probe(device, FASTMATH_FN[fn], x.size, x.ctypes.data_as(_dp), y0.ctypes.data_as(_dp),
                                   y1.ctypes.data_as(_dp)))
    return y0, y1


def env_dims(env_id):
    v = [C.c_int32() for _ in range(6)]
    check(lib().i2c_env_dims(env_id, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)  # dx, du, dz, dzt, n_par, dy
