"""Sub-warp-per-problem kernel (csrc/i2c_group.cuh: G lanes cooperate on one problem) against the per-thread kernels and
the oracle.  I2C_B200_GROUP=1/0 forces / forbids the variant (the launcher otherwise picks it for small batches)."""
import os

import numpy as np
import pytest

from conftest import GAINS, golden, relerr
from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu

FIELDS = ["mu_xu0_f", "sig_xu0_f", "mu_z0_f", "sig_z0_f", "mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn", "mu_x3_m",
          "sig_x3_m", "mu_xu0_m", "sig_xu0_m", "mu_z0_m", "sig_z0_m", "K", "k", "sigK"]
PF = ["mu_xu0_pf", "sig_xu0_pf", "mu_z0_pf", "sig_z0_pf", "mu_x3_pf", "sig_x3_pf"]


class mode:
    def __init__(self, group):
        self.v = str(int(group))

    def __enter__(self):
        self.old = os.environ.get("I2C_B200_GROUP")
        os.environ["I2C_B200_GROUP"] = self.v

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("I2C_B200_GROUP", None)
        else:
            os.environ["I2C_B200_GROUP"] = self.old


CASES = [
    # tolerances: see tests/test_gpu_parity.py (max(1e-9, ~10 x measured); gains per environment)
    ("PendulumKnown", 70, 40, np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, 0.0, [0.3, 0.5], 2.0, 1e-9, 2e-9),
    ("CartpoleKnown", 37, 30, np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]), 80.0, 0.0, 0.05, 1.0, 1e-9, 2e-8),
    ("DoubleCartpoleKnown", 33, 30, 1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.99,
     0.02, 1.0, 1e-9, 5e-8),
    ("LinearKnownMinimumEnergy", 9, 20, None, np.diag([1.0]), 10.0, 0.5, 0.3, 10.0, 1e-10, 5e-8),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_group_kernel_vs_thread_kernel_and_oracle(i2c_b200, case):
    from oracle import i2c_oracle as O

    env, B, T, Q, R, alpha, tol, xs, su, tol_s, tol_g = case
    e = i2c_b200.envs.make(env)
    Qf = Q if e.has_term else None
    rng = np.random.default_rng(12)
    x0 = e.x0 + np.asarray(xs) * rng.normal(size=(B, e.dim_x))
    mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
    args = (env, B, T, Q, R, Qf, alpha, tol, mu_u, su * np.eye(e.dim_u))
    Gt = i2c_b200.BatchedI2c(*args, x0=x0, enable_aux=True)
    Gg = i2c_b200.BatchedI2c(*args, x0=x0, enable_aux=True)
    ref = O.make_graph(env, T, Q, R, Qf, alpha, tol, mu_u, su * np.eye(e.dim_u), B=B, x0=x0)
    for it in range(3):
        with mode(0):
            Gt.learn(1)
        with mode(1):
            Gg.learn(1)
        ref.learn_msgs()
        assert np.all(Gg.status()[0] == 0), (it, Gg.status())
        for a in FIELDS:
            fl = 1e-6 if a in GAINS else 0.0
            e1 = relerr(Gg.field(a), Gt.field(a), floor=fl)
            assert e1 < (tol_g if a in GAINS else tol_s), (it, a, "vs thread", e1)
        for a in ["mu_xu1_f", "sig_xu1_f", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK", "mu_z0_m", "sig_z0_m"]:
            e2 = relerr(Gg.field(a), ref.stack(a), floor=1e-6 if a in GAINS else 0.0)
            assert e2 < (tol_g if a in GAINS else tol_s), (it, a, "vs oracle", e2)
        assert relerr(Gg.alpha, Gt.alpha) < 1e-11 and relerr(Gg.alpha, ref.alpha) < 1e-10
    for name in ["alpha", "alpha_desired", "cost_m", "cost_m_var", "policy_entropy", "x_prior_entropy"]:
        assert relerr(np.array(Gg.metrics[name]), np.array(Gt.metrics[name])) < 1e-10, name
    # the variants share records: alternate them between launches
    with mode(0):
        Gg.learn(1)
        Gt.learn(1)
    assert relerr(Gg.field("K"), Gt.field("K"), floor=1e-6) < tol_g


def test_group_kernel_covariance_control_propagate_golden(i2c_b200):
    """Config-4 shaped run (double cart-pole, covariance control, propagate) against the reference golden."""
    capi = i2c_b200.capi
    g = golden("double_cartpole_covctrl_T50")
    G = i2c_b200.BatchedI2c(str(g["env"]), 1, int(g["T"]), g["Q"], g["R"], g["Qf"], float(g["alpha0"]), float(g["tol"]),
                            g["mu_u"], g["sig_u"], g["mu_x_term"], g["sig_x_term"], x0=g["x0"], enable_aux=True)
    with mode(1):
        G._propagate = True
        G.set_cell_flag(capi.CELL_EXPERT, bool(g["expert"]))
        G.run(1, capi.PH_PROPAGATE, False)
        for a in PF:
            assert relerr(G.field(a)[0], g[f"it0/{a}"]) < 1e-9, a
        n_dump, n_total = int(g["n_dump"]), int(g["n_total"])
        for it in range(1, n_total + 1):
            G.learn(1)
            assert np.all(G.status()[0] == 0), G.status()
            if it <= n_dump:
                for a in FIELDS[2:] + PF:
                    e = relerr(G.field(a)[0], g[f"it{it}/{a}"], floor=1e-6 if a in GAINS else 0.0)
                    assert e < (5e-8 if a in GAINS else 3e-9), (it, a, e)  # measured 4.5e-9 / 3e-10
    assert relerr(np.array([a[0] for a in G.alphas]), g["alphas"]) < 1e-9
    assert relerr(np.array(G.metrics["cost_m"])[:, 0], g["costs_m"]) < 1e-9
    assert relerr(np.array(G.metrics["cost_pf"])[:, 0], g["costs_pf"]) < 1e-9
    assert relerr(np.array(G.metrics["kl_term"])[:, 0], g["kl_terms"]) < 1e-9


def test_group_kernel_expert_propagate_golden(i2c_b200):
    capi = i2c_b200.capi
    g = golden("pendulum_propagate_expert_T50")
    G = i2c_b200.BatchedI2c(str(g["env"]), 1, int(g["T"]), g["Q"], g["R"], g["Qf"], float(g["alpha0"]), float(g["tol"]),
                            g["mu_u"], g["sig_u"], x0=g["x0"], enable_aux=True)
    with mode(1):
        G._propagate = True
        G.set_cell_flag(capi.CELL_EXPERT, True)
        G.run(1, capi.PH_PROPAGATE, False)
        for it in range(1, int(g["n_total"]) + 1):
            G.learn(1)
            if it <= int(g["n_dump"]):
                for a in FIELDS[2:] + PF:
                    e = relerr(G.field(a)[0], g[f"it{it}/{a}"], floor=1e-6 if a in GAINS else 0.0)
                    assert e < (3e-9 if a in GAINS else 1e-9), (it, a, e)  # measured 2.4e-10
    assert relerr(np.array([a[0] for a in G.alphas]), g["alphas"]) < 1e-9
    assert relerr(np.array(G.metrics["cost_pf"])[:, 0], g["costs_pf"]) < 1e-9


def test_group_kernel_timing_report(i2c_b200):
    """Measurement only (printed with -s): per-GPU shards of BASELINE configs 4 (2048 double cart-pole problems) and 3.
    Round-1 status: the variant is parity-complete but NOT yet faster than the per-thread kernels (its matrices live in
    shared memory and every small factorisation pays an exchange per column: see DESIGN.md), so the launcher keeps it
    opt-in (I2C_B200_GROUP=1)."""
    rng = np.random.default_rng(0)
    out = {}
    for env, B, T, Q, R, alpha, tol in [
        ("DoubleCartpoleKnown", 2048, 100, 1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.99),
        ("PendulumKnown", 4096, 200, np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, 0.0),
    ]:
        e = i2c_b200.envs.make(env)
        x0 = e.x0 + 0.02 * rng.normal(size=(B, e.dim_x))
        mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
        for grp in (0, 1):
            G = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, alpha, tol, mu_u, np.eye(e.dim_u), x0=x0, max_iters=8)
            with mode(grp):
                G.learn(2, collect=False)
                G.synchronize()
                G.learn(3, collect=False)
                G.synchronize()
            out[(env, grp)] = G.last_run_ms() / 3
            assert np.all(G.status()[0] == 0)
            G.close()
    print({f"{k[0]} group={k[1]}": round(v, 3) for k, v in out.items()})
