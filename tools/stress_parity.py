#!/usr/bin/env python
"""Randomised parity sweep (not part of the test-suite): many seeds x environments x kernel variants against the oracle;
prints the worst norm-wise error per field.  python tools/stress_parity.py [n_seeds]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "input-inference-for-control_b200"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import i2c_b200  # noqa: E402
from conftest import relerr  # noqa: E402
from oracle import i2c_oracle as O  # noqa: E402

CASES = {
    "PendulumKnown": (np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, 0.0, [0.3, 0.5], 2.0, 40),
    "CartpoleKnown": (np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]), 80.0, 0.0, 0.05, 1.0, 30),
    "DoubleCartpoleKnown": (1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.99, 0.02, 1.0, 25),
}
FIELDS = ["mu_xu1_f", "sig_xu1_f", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK"]
GAINS = ("K", "k")
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
worst = {}
for env, (Q, R, alpha, tol, xs, su, T) in CASES.items():
    e = i2c_b200.envs.make(env)
    for seed in range(n_seeds):
        rng = np.random.default_rng(1000 + seed)
        B = int(rng.integers(1, 70))
        x0 = e.x0 + np.asarray(xs) * rng.normal(size=(B, e.dim_x))
        mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
        ref = O.make_graph(env, T, Q, R, Q, alpha, tol, mu_u, su * np.eye(e.dim_u), B=B, x0=x0)
        for _ in range(3):
            ref.learn_msgs()
        for grp in ("0", "1"):
            os.environ["I2C_B200_GROUP"] = grp
            G = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, alpha, tol, mu_u, su * np.eye(e.dim_u), x0=x0, enable_aux=True)
            G.learn(3)
            ok = bool(np.all(G.status()[0] == 0))
            for f in FIELDS:
                err = relerr(G.field(f), ref.stack(f), floor=1e-6 if f in GAINS else 0.0)
                key = (env, grp, f)
                worst[key] = max(worst.get(key, 0.0), err if ok else float("inf"))
            key = (env, grp, "alpha")
            worst[key] = max(worst.get(key, 0.0), relerr(G.alpha, ref.alpha))
            G.close()
for (env, grp, f), v in sorted(worst.items()):
    print(f"{env:22s} group={grp} {f:10s} {v:.2e}")
