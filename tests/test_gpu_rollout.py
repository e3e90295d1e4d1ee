"""batched policy-evaluation roll-outs (i2c/env.py:40-103) on the GPU vs the NumPy oracle with
host-fed disturbances, for controllers produced by the CUDA EM sweep."""
import numpy as np
import pytest

from conftest import relerr
from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env,Q,R,alpha,xs", [
    ("PendulumKnown", np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, [0.3, 0.5]),
    ("CartpoleKnown", np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]), 80.0, 0.05),
])
def test_rollout_vs_oracle(i2c_b200, env, Q, R, alpha, xs):
    from oracle import envs as E
    from oracle.rollout import rollout as ref_rollout

    rng = np.random.default_rng(0)
    e = i2c_b200.envs.make(env)
    B, T, Rr = 6, 40, 5
    x0 = e.x0 + np.asarray(xs) * rng.normal(size=(B, e.dim_x))
    mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
    G = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, alpha, 0.0, mu_u, np.eye(e.dim_u), x0=x0)
    G.learn(5)
    K, k, sk = G.get_local_linear_policy()
    mu = G.field("mu_xu0_m")[:, :, : e.dim_x]
    lam = np.linalg.inv(G.field("sig_xu0_m")[:, :, : e.dim_x, : e.dim_x])
    sys_ = E.make(env)
    x_init = np.repeat(x0[:, None, :], Rr, axis=1)
    Ls = np.linalg.cholesky(e.sig_eta)
    eta = rng.normal(size=(B, Rr, T, e.dim_x)) @ Ls.T
    eps_u = rng.normal(size=(B, Rr, T, e.dim_u))
    # plain linear policy, deterministic action (i2c_run.py:96-100)
    xu, z, zt = i2c_b200.rollout(env, x_init, K, k, eta=eta)
    xr, zr, ztr = ref_rollout(sys_, x_init, K, k, eta)
    assert relerr(xu, xr) < 1e-9 and relerr(z, zr) < 1e-9 and relerr(zt, ztr) < 1e-9
    # stochastic actions
    xu, z, zt = i2c_b200.rollout(env, x_init, K, k, sig_k=sk, eta=eta, eps_u=eps_u)
    xr, zr, ztr = ref_rollout(sys_, x_init, K, k, eta, sig_k=sk, eps_u=eps_u)
    assert relerr(xu, xr) < 1e-9 and relerr(zt, ztr) < 1e-9
    # expert (gated) policy, soft and hard (i2c_run.py:102-104; policy/linear.py:73-90); k = mu_u0_m there
    k_exp = G.field("mu_xu0_m")[:, :, e.dim_x:]
    for soft in (True, False):
        xu, z, zt = i2c_b200.rollout(env, x_init, K, k_exp, expert=(mu, lam), soft_expert=soft, eta=eta)
        xr, zr, ztr = ref_rollout(sys_, x_init, K, k_exp, eta, expert=(mu, lam), soft=soft)
        assert relerr(xu, xr) < 1e-8 and relerr(zt, ztr) < 1e-8, soft


def test_rollout_device_rng_statistics(i2c_b200):
    """Device-side Philox disturbances: zero-gain policy on the pendulum start state; the one-step spread of many
    roll-outs matches sig_eta, runs are reproducible per seed and differ across seeds."""
    env = "PendulumKnown"
    e = i2c_b200.envs.make(env)
    B, Rr, T = 2, 4096, 3
    x_init = np.broadcast_to(e.x0, (B, Rr, 2)).copy()
    K, k = np.zeros((B, T, 1, 2)), np.zeros((B, T, 1))
    a = i2c_b200.rollout(env, x_init, K, k, seed=1)[0]
    b = i2c_b200.rollout(env, x_init, K, k, seed=1)[0]
    c = i2c_b200.rollout(env, x_init, K, k, seed=2)[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    step = a[:, :, 1, :2]  # state after one noisy step
    cov = np.cov(step.reshape(-1, 2).T)
    assert np.allclose(np.diag(cov), np.diag(e.sig_eta), rtol=0.1)
    nf = i2c_b200.rollout(env, x_init, K, k, sig_eta=np.zeros((2, 2)))[0]
    assert np.ptp(nf[:, :, 1, 0]) == 0.0


@pytest.mark.parametrize("tag", ["det_env_det_pol", "noisy_env_det_pol", "noisy_env_noisy_pol", "expert_soft", "expert_hard"])
@pytest.mark.parametrize("env", ["PendulumKnown", "CartpoleKnown"])
def test_rollout_vs_reference_sim_golden(i2c_b200, env, tag):
    """The CUDA roll-out against the reference's own BaseSim.run (rollout_kat.npz: seeded global RNG, realised
    disturbances logged and fed to the kernel)."""
    from conftest import golden
    from test_oracle_rollout import golden_case

    g = golden("rollout_kat")
    T, x_init, K, k, eta, kw = golden_case(g, env, tag)
    if "soft" in kw:
        kw["soft_expert"] = kw.pop("soft")
    xu, z, zt = i2c_b200.rollout(env, x_init, K, k, eta=eta, **kw)
    assert relerr(xu[0, 0], g[f"{env}/{tag}/xt"]) < 1e-10
    assert relerr(z[0, 0], g[f"{env}/{tag}/zt"]) < 1e-10
    assert relerr(zt[0, 0], g[f"{env}/{tag}/z_term"].reshape(-1)) < 1e-10
