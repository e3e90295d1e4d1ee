import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "input-inference-for-control_b200"))
import __graft_entry__ as ge
ge.build()
import i2c_b200
from i2c_b200 import capi
sf = 1e-3
Q = sf * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]); R = sf * np.diag([0.1])
mu_t, sig_t = np.zeros(6), np.diag([0.01, 0.005, 0.005, 0.05, 0.05, 0.05])
for B, T, expert in [(64, 500, False), (64, 100, False), (64, 200, False), (64, 500, True)]:
    rng = np.random.default_rng(4321)
    e = i2c_b200.envs.make("DoubleCartpoleKnown")
    x0 = e.x0 + 0.05 * rng.normal(size=(B, 6)); mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    for grp in ("0", "1"):
        os.environ["I2C_B200_GROUP"] = grp
        G = i2c_b200.BatchedI2c("DoubleCartpoleKnown", B, T, Q, R, Q, 0.05, 0.99, mu_u, np.eye(1), mu_t, sig_t, x0=x0, enable_aux=True)
        G._propagate = True
        G.set_cell_flag(capi.CELL_EXPERT, expert)
        G.propagate()
        st, info = G.status()
        print(B, T, expert, "group", grp, "status", np.unique(st, return_counts=True), "cell", np.unique(info & 0xffff)[:5],
              "max sig_x3_pf", float(np.nanmax(np.abs(G.field("sig_x3_pf")[:, -1]))))
        G.close()
