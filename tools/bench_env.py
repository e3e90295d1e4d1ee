#!/usr/bin/env python
"""Quick per-environment throughput probe of the persistent EM kernel (not the judged bench): prints updates/s and
roofline fractions for a given env / batch / horizon using the SURVEY.md 8(d) algorithmic figures."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "input-inference-for-control_b200"))

ALG = {  # env: (flops, bytes) per problem-timestep update
    "LinearKnown": (2327, 696), "PendulumKnown": (3192, 704), "CartpoleKnown": (11601, 1792),
    "DoubleCartpoleKnown": (34473, 3400), "Quadrotor": (33105, 4160)}
HYP = {
    "PendulumKnown": dict(Q=np.diag([1.0, 100.0, 1.0]), R=np.diag([2.0]), alpha=100.0, tol=0.0, sig_u=2.0, xs=[0.3, 0.5]),
    "CartpoleKnown": dict(Q=np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), R=np.diag([1.0]), alpha=80.0, tol=0.0, sig_u=1.0, xs=0.05),
    "DoubleCartpoleKnown": dict(Q=1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), R=1e-4 * np.eye(1),
                                alpha=0.05, tol=0.99, sig_u=1.0, xs=0.02),
    "Quadrotor": dict(Q=np.diag([1e3, 1e3, 1e3, 1, 1, 1]), R=np.diag([1e-3, 1e-3]), alpha=1.0, tol=1.0, sig_u=1e-2, xs=0.01),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="PendulumKnown")
    ap.add_argument("--problems", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--propagate", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    import i2c_b200
    from i2c_b200 import capi

    h = HYP[a.env]
    e = i2c_b200.envs.make(a.env)
    rng = np.random.default_rng(0)
    x0 = e.x0 + np.asarray(h["xs"]) * rng.normal(size=(a.problems, e.dim_x))
    mu0 = 0.5 * 9.81 * i2c_b200.envs.QUAD_MASS if a.env == "Quadrotor" else 0.0
    mu_u = mu0 + 1e-2 * rng.normal(size=(a.problems, a.horizon, e.dim_u))
    Qf = h["Q"] / (1e3 if a.env == "Quadrotor" else 1.0)
    g = i2c_b200.BatchedI2c(a.env, a.problems, a.horizon, h["Q"], h["R"], Qf, h["alpha"], h["tol"], mu_u,
                            h["sig_u"] * np.eye(e.dim_u), x0=x0, max_iters=max(a.iters, 3))
    ph = capi.PH_LEARN | (capi.PH_PROPAGATE if a.propagate else 0)
    g.run(2, ph, collect=False)
    g.run(a.iters, ph, collect=False)
    g.synchronize()
    ms = g.last_run_ms()
    rate = a.problems * a.horizon * a.iters / (ms * 1e-3)
    F, Bb = ALG[a.env]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    fp64 = capi.dfma_peak(0)
    print(json.dumps({"env": a.env, "problems": a.problems, "horizon": a.horizon, "iters": a.iters, "ms_per_iter": ms / a.iters,
                      "updates_per_s": rate, "hbm_frac": Bb * rate / 1e9 / hbm, "fp64_frac_measured": F * rate / 1e12 / fp64,
                      "failed": int(np.count_nonzero(g.status()[0]))}))


if __name__ == "__main__":
    main()
