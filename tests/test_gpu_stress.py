"""Randomised parity sweep (the former tools/stress_parity.py, now part of `-m gpu`): seeds x environments x both kernel
families against the oracle, long horizons (T = 200 / 120 / 60) and 10 EM iterations, random batch sizes.  Tolerances:
max(1e-9, ~10 x measured) per environment (profiles/r02_parity_floors.txt): ten alpha-coupled EM iterations of the pendulum
swing-up over T = 200 amplify round-off to 3e-9 on the states (cart-pole / double cart-pole stay at 1e-11 / 4e-11)."""
import os

import numpy as np
import pytest

from conftest import relerr
from test_gpu_parity import i2c_b200  # noqa: F401
from tools_inputs import HYP

pytestmark = pytest.mark.gpu

FIELDS = ["mu_xu1_f", "sig_xu1_f", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK"]
GAINS = ("K", "k")
HORIZON = {"PendulumKnown": 200, "CartpoleKnown": 120, "DoubleCartpoleKnown": 60}
ITERS = 10
# measured worst errors over the sweep: states 3.2e-9 / 7e-12 / 4e-11, gains 1.7e-9 / 4.3e-10 / 4.9e-9
TOL_STATE = {"PendulumKnown": 3e-8, "CartpoleKnown": 1e-9, "DoubleCartpoleKnown": 1e-9}
TOL_GAINS = {"PendulumKnown": 2e-8, "CartpoleKnown": 5e-9, "DoubleCartpoleKnown": 5e-8}


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("env", list(HORIZON))
def test_random_batches_long_horizon(i2c_b200, env, seed):
    from oracle import i2c_oracle as O

    h, T = HYP[env], HORIZON[env]
    e = i2c_b200.envs.make(env)
    rng = np.random.default_rng(2000 + seed)
    B = int(rng.integers(1, 48))
    x0 = e.x0 + np.asarray(h["xs"]) * rng.normal(size=(B, e.dim_x))
    mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
    su = h["sig_u"] * np.eye(e.dim_u)
    ref = O.make_graph(env, T, h["Q"], h["R"], h["Q"], h["alpha"], h["tol"], mu_u, su, B=B, x0=x0)
    for _ in range(ITERS):
        ref.learn_msgs()
    for grp in ("0", "1"):
        old = os.environ.get("I2C_B200_GROUP")
        os.environ["I2C_B200_GROUP"] = grp
        try:
            G = i2c_b200.BatchedI2c(env, B, T, h["Q"], h["R"], h["Q"], h["alpha"], h["tol"], mu_u, su, x0=x0, enable_aux=(seed % 2 == 0),
                                    max_iters=ITERS)
            G.learn(ITERS)
        finally:
            if old is None:
                os.environ.pop("I2C_B200_GROUP", None)
            else:
                os.environ["I2C_B200_GROUP"] = old
        assert np.all(G.status()[0] == 0), (grp, G.status())
        for f in FIELDS:
            err = relerr(G.field(f), ref.stack(f), floor=1e-6 if f in GAINS else 0.0)
            tol = TOL_GAINS[env] if f in GAINS else TOL_STATE[env]
            assert err < tol, (env, seed, grp, f, err)
        assert relerr(G.alpha, ref.alpha) < 1e-9
        assert relerr(np.array(G.metrics["cost_m"]), np.array(ref.costs_m)) < 1e-8
        G.close()
