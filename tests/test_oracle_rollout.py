"""Pins oracle/rollout.py to the reference: BaseSim.run (i2c/env.py:40-74) run with a seeded global RNG through the
unmodified reference, every disturbance logged (tests/golden/make_golden.py: rollouts -> rollout_kat.npz)."""
import numpy as np
import pytest

from conftest import golden, relerr

ENVS = ["PendulumKnown", "CartpoleKnown"]
RUNS = ["det_env_det_pol", "noisy_env_det_pol", "noisy_env_noisy_pol", "expert_soft", "expert_hard"]


def golden_case(g, env, tag):
    """Arguments of oracle.rollout.rollout / i2c_b200.rollout for one golden run (B = R = 1)."""
    T = int(g[f"{env}/T"])
    expert = tag.startswith("expert")
    K = g[f"{env}/ex_K" if expert else f"{env}/K"][None]
    k = g[f"{env}/ex_k" if expert else f"{env}/k"][None]
    sk = g[f"{env}/sigK"][None]
    x_init = g[f"{env}/x0"][None, None, :]
    eta = g[f"{env}/{tag}/eta"][None, None]
    kw = {}
    if tag == "noisy_env_noisy_pol":
        Ls = np.linalg.cholesky(sk[0])  # [T, du, du]
        eps = np.linalg.solve(Ls, g[f"{env}/{tag}/u_noise"][:, :, None])[:, :, 0]
        kw = dict(sig_k=sk, eps_u=eps[None, None])
    if expert:
        kw = dict(expert=(g[f"{env}/ex_mu"][None], g[f"{env}/ex_lam"][None]), soft=tag.endswith("soft"))
    return T, x_init, K, k, eta, kw


@pytest.mark.parametrize("tag", RUNS)
@pytest.mark.parametrize("env", ENVS)
def test_oracle_rollout_reproduces_reference_sim(env, tag):
    from oracle import envs as E
    from oracle.rollout import rollout

    g = golden("rollout_kat")
    T, x_init, K, k, eta, kw = golden_case(g, env, tag)
    xu, z, zt = rollout(E.make(env), x_init, K, k, eta, **kw)
    assert relerr(xu[0, 0], g[f"{env}/{tag}/xt"]) < 1e-12
    assert relerr(z[0, 0], g[f"{env}/{tag}/zt"]) < 1e-12
    assert relerr(zt[0, 0], g[f"{env}/{tag}/z_term"].reshape(-1)) < 1e-12
