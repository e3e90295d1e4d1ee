"""All kernel variants compute the same numbers: the launcher picks the team kernel (<= 296 tiles), the cp.async
latency variant, the TMA-bulk throughput variant or the 128-register throughput variant from the batch size; run the
same seeded problems through each regime (by chunking the batch) and against the oracle on a subset.  Also covers
size-independent properties at the full BASELINE batch size."""
import numpy as np
import pytest

from conftest import relerr
from test_gpu_parity import TOL_GAIN, i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def inputs(B, T, seed=3):
    rng = np.random.default_rng(seed)
    x0 = np.array([np.pi, 0.0]) + np.array([0.3, 0.5]) * rng.normal(size=(B, 2))
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    return x0, mu_u


Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])


def run(m, x0, mu_u, iters):
    B, T = mu_u.shape[:2]
    g = m.BatchedI2c("PendulumKnown", B, T, Q, R, Q, 100.0, 0.0, mu_u, 2.0 * np.eye(1), x0=x0)
    g.learn(iters)
    assert np.all(g.status()[0] == 0)
    K, k, s = g.get_local_linear_policy()
    return dict(K=K, k=k, sigK=s, mu=g.field("mu_xu0_m"), sig=g.field("sig_xu0_m"), alpha=g.alpha,
                cost=np.array(g.metrics["cost_m"]))


@pytest.mark.parametrize("B,T", [(60000, 10), (20000, 16), (6000, 24)])
def test_regimes_agree(i2c_b200, B, T):
    """B = 60000 -> em_kernel<.,4> (128 registers, TMA bulk); 20000 -> em_kernel<.,1> bulk; 6000 -> cp.async variant;
    chunks of 4096 -> team kernel."""
    x0, mu_u = inputs(B, T)
    full = run(i2c_b200, x0, mu_u, 3)
    sub = slice(4096, 8192) if B >= 8192 else slice(0, 4096)
    part = run(i2c_b200, x0[sub], mu_u[sub], 3)
    for key in ("mu", "sig", "sigK", "alpha"):
        assert relerr(full[key][sub], part[key]) < 1e-11, key
    for key in ("K", "k"):
        assert relerr(full[key][sub], part[key], 1e-6) < 1e-9, key
    assert relerr(full["cost"][:, sub], part["cost"]) < 1e-11


def test_full_size_against_oracle_subset(i2c_b200):
    """BASELINE config 3 at full size (4096 x T=200): a random subset of problems against the oracle."""
    from oracle import i2c_oracle as O

    B, T, iters = 4096, 200, 3
    x0, mu_u = inputs(B, T, seed=1234)
    full = run(i2c_b200, x0, mu_u, iters)
    idx = np.random.default_rng(0).choice(B, 24, replace=False)
    ref = O.make_graph("PendulumKnown", T, Q, R, Q, 100.0, 0.0, mu_u[idx], 2.0 * np.eye(1), B=len(idx), x0=x0[idx])
    for _ in range(iters):
        ref.learn_msgs()
    assert relerr(full["mu"][idx], ref.stack("mu_xu0_m")) < 1e-9
    assert relerr(full["sig"][idx], ref.stack("sig_xu0_m")) < 1e-9
    assert relerr(full["alpha"][idx], ref.alpha) < 1e-9  # round-off amplified over 3 sweeps x 200 cells
    Kr, kr, sr = ref.get_local_linear_policy()
    assert relerr(full["K"][idx], Kr, 1e-6) < 2e-8 and relerr(full["k"][idx], kr, 1e-6) < 2e-8  # T = 200: measured 3e-9
    # size-independent properties over the whole batch: PD posteriors, sigK > 0, finite costs, alpha > 0
    assert np.all(np.isfinite(full["cost"])) and np.all(full["alpha"] > 0) and np.all(full["sigK"] > 0)
    assert np.all(np.linalg.eigvalsh(full["sig"].reshape(-1, 3, 3)) > 0)
    # permutation equivariance: problems are independent
    perm = np.random.default_rng(1).permutation(B)
    again = run(i2c_b200, x0[perm], mu_u[perm], iters)
    assert np.array_equal(again["K"], full["K"][perm]) and np.array_equal(again["alpha"], full["alpha"][perm])


@pytest.mark.parametrize("env,B,T,kw", [
    ("PendulumKnown", 200, 37, {}),                      # team kernel, 8 warps: copy warp + 6 tail warps
    ("PendulumKnown", 6000, 12, {}),                     # team kernel, 4 warps (two blocks per SM)
    ("CartpoleKnown", 96, 25, {}),                       # records read in place (no register copy of the ring slot)
    ("PendulumKnown", 130, 30, {"propagate": True}),     # propagate sweep through the record ring
    ("Quadrotor", 64, 12, {}),                           # HOT without a staging ring (records read from global memory)
])
def test_hot_specialisation_equals_generic(i2c_b200, monkeypatch, env, B, T, kw):
    """The HOT team kernel (common configuration compiled in, copy warp + record ring, plain-sweep loop) computes what the
    generic team kernel computes (same arithmetic apart from operation order: means / covariances to 1e-11, gains -- solves
    against 1e-5 covariances -- to the per-environment gain tolerance of test_gpu_parity.py; measured 1.2e-10 / 3.2e-9); the first EM iteration has independent cells (generic loop),
    the later ones take the plain loop."""
    from i2c_b200 import capi
    from tools_inputs import make_case  # noqa: F401  (tests/tools_inputs.py)

    outs = []
    for no_hot in (False, True):
        if no_hot:
            monkeypatch.setenv("I2C_B200_NO_HOT", "1")
        else:
            monkeypatch.delenv("I2C_B200_NO_HOT", raising=False)
        g = make_case(i2c_b200, env, B, T)
        ph = capi.PH_LEARN | (capi.PH_PROPAGATE if kw.get("propagate") else 0)
        for _ in range(3):
            g.run(1, ph)
        assert np.all(g.status()[0] == 0)
        K, k, s = g.get_local_linear_policy()
        outs.append(dict(K=K, k=k, sigK=s, mu=g.field("mu_xu0_m"), sig=g.field("sig_xu0_m"), alpha=g.alpha,
                         cost=np.array(g.metrics["cost_m"]), J=g.field("J_dyn"), f=g.field("sig_xu1_f")))
    a, b = outs
    for key in a:
        assert relerr(a[key], b[key], 1e-9) < (TOL_GAIN.get(env, 2e-8) if key in ("K", "k", "J") else 1e-11), key


@pytest.mark.parametrize("kw", [{}, {"propagate": True}, {"fb_only": True}, {"covctrl": True}])
def test_ticket_kernel_equals_static_wave(i2c_b200, kw, monkeypatch):
    """em_ticket_kernel ((tile, iteration) work items drawn from a ticket counter, loop-carried state rebuilt per item) is
    bit-identical to the static one-warp-per-tile launch of the same 168-register variant: learn_msgs with and without the
    in-loop propagate, and the MPC optimise loop (forward + backward + _update_priors, no M-step)."""
    B, T, iters = 3000, 20, 5
    x0, mu_u = inputs(B, T, seed=11)
    out = {}
    for minb in ("3", "5"):
        monkeypatch.setenv("I2C_B200_MINB", minb)
        if kw.get("covctrl"):
            # covariance control (pendulum_known_act_reg_quad.py:22-33): terminal distribution, temperature schedule carried
            # across the iterations, in-loop propagate, KL metric
            g = i2c_b200.BatchedI2c("PendulumKnownActReg", B, T, None, np.diag([1.0]), None, 300.0, 1.0, 0.0 * mu_u,
                                    0.5 * np.eye(1), np.array([0.0, 0.0]), np.diag([1e-3, 1e-3]),
                                    x0=np.array([np.pi, 0.0]) + 0.05 * (x0 - np.array([np.pi, 0.0])))
            g._propagate = True
            g.set_cell_flag(i2c_b200.capi.CELL_EXPERT, False)
        else:
            g = i2c_b200.BatchedI2c("PendulumKnown", B, T, Q, R, Q, 100.0, 0.0, mu_u, 2.0 * np.eye(1), x0=x0)
            g._propagate = bool(kw.get("propagate"))
        if kw.get("fb_only"):
            g.tau = T
            g.forward_backward(iters, update_priors=True)
            g.forward_backward(2, update_priors=True)  # a second launch continues from the swapped record roles
        else:
            g.learn(iters)
            g.learn(2)
        if not kw.get("covctrl"):
            assert np.all(g.status()[0] == 0)
        K, k, s = g.get_local_linear_policy()
        out[minb] = dict(K=K, k=k, sigK=s, mu=g.field("mu_xu0_m"), sig=g.field("sig_xu0_m"), alpha_final=g.alpha,
                         status=g.status()[0], temp=np.array(g.temp),
                         **{m: np.array(v) for m, v in g.metrics.items()})
    for key in out["3"]:
        assert np.array_equal(out["3"][key], out["5"][key], equal_nan=True), key
