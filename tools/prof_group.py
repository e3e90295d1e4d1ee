import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "input-inference-for-control_b200"))
import __graft_entry__ as ge
ge.build()
import i2c_b200
env, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
e = i2c_b200.envs.make(env)
if env == "DoubleCartpoleKnown":
    Q, R, alpha, tol = 1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.99
else:
    Q, R, alpha, tol = np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, 0.0
x0 = e.x0 + 0.02 * rng.normal(size=(B, e.dim_x))
mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
G = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, alpha, tol, mu_u, np.eye(e.dim_u), x0=x0, max_iters=8)
G.learn(2, collect=False); G.synchronize()
G.learn(2, collect=False); G.synchronize()
print(G.last_run_ms() / 2, "ms/iter")
