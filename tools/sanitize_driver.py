#!/usr/bin/env python
"""Small launches of every hot kernel family, to be run under compute-sanitizer (tools/sanitize.sh)."""
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
from i2c_b200 import capi  # noqa: E402
from tools_inputs import make_case  # noqa: E402


def run(env, B, T, iters=2, propagate=False, **env_vars):
    old = {k: os.environ.get(k) for k in env_vars}
    os.environ.update({k: str(v) for k, v in env_vars.items()})
    try:
        g = make_case(i2c_b200, env, B, T)
        ph = capi.PH_LEARN | (capi.PH_PROPAGATE if propagate else 0)
        g.run(iters, ph)
        g.synchronize()
        ok = int(np.count_nonzero(g.status()[0])) == 0
        print(f"{env:22s} B={B:5d} T={T:3d} {env_vars} ok={ok}", flush=True)
        g.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


run("PendulumKnown", 96, 24)                                  # em_team_kernel<.,8,HOT>: copy warp + ring + plain loop
run("PendulumKnown", 96, 24, propagate=True)                  # ... propagate sweep through the ring
run("PendulumKnown", 96, 24, I2C_B200_NO_HOT=1)               # em_team_kernel<.,8> generic (cp.async stream)
run("PendulumKnown", 160 * 32, 10)                            # em_team_kernel<.,4,HOT> (two blocks per SM)
run("CartpoleKnown", 64, 16)                                  # HOT, records read in place
run("Quadrotor", 64, 10)                                      # HOT without staging
run("PendulumKnown", 300 * 32, 8)                             # em_kernel<.,1,LAT> cp.async
run("PendulumKnown", 700 * 32, 6)                             # em_kernel<.,1> TMA bulk
run("PendulumKnown", 148 * 12 * 32, 4)                        # em_kernel<.,3/4> throughput variants
run("PendulumKnown", 200 * 32, 5, iters=4, I2C_B200_MINB=5)   # em_ticket_kernel: (tile, iteration) work items, gpu-scope release / acquire
run("PendulumKnown", 64 * 32, 6, iters=3, propagate=True, I2C_B200_MINB=5)
run("DoubleCartpoleKnown", 64, 12)                            # em_group_kernel<.,8>
run("DoubleCartpoleKnown", 64, 12, I2C_B200_GROUP=0)          # per-thread kernel for the large system
# parallel-in-time scan
rng = np.random.default_rng(5)
B, T = 32, 256
A = np.array([[1.0, 0.1], [-0.05, 0.98]]) + 0.01 * rng.normal(size=(B, 2, 2))
xg = rng.normal(size=(B, 2))
par = i2c_b200.envs.linear_params(A, np.array([[0.0], [0.1]]), xg - np.einsum("bij,bj->bi", A, xg))
z = np.repeat(np.concatenate((xg, np.zeros((B, 1))), axis=1)[:, None, :], T, axis=1)
G = i2c_b200.BatchedI2c("LinearKnown", B, T, np.diag([1.0, 2.0]), np.diag([0.5]), np.diag([1.0, 2.0]), 5.0, 0.5,
                        np.zeros((B, T, 1)), np.eye(1), x0=xg + 2.0, sig_x0=1e-2 * np.eye(2), sig_eta=1e-3 * np.eye(2),
                        env_par=par, z=z, z_term=xg, z_per_problem=True, inference="linearize", max_iters=4)
G.time_parallel_chunk = 32
G.forward_backward(2)
G.synchronize()
print("scan ok", bool(np.all(G.status()[0] == 0)), flush=True)
# stand-alone transforms and roll-outs
m = rng.normal(size=(40, 3))
S = np.einsum("bij,bkj->bik", *(2 * [0.1 * rng.normal(size=(40, 3, 3))])) + 1e-2 * np.eye(3)
i2c_b200.quadrature("PendulumKnown", "observe", m, S)
i2c_b200.quadrature("PendulumKnown", "forward", m, S)
K, k = np.zeros((2, 5, 1, 2)), np.zeros((2, 5, 1))
i2c_b200.rollout("PendulumKnown", np.zeros((2, 8, 2)), K, k, seed=1)
# closed-loop MPC step: CKF kernel, HOT = 2 team kernel (cells with their own alpha), tail kernel (first action + horizon shift),
# page-locked action ring; pipelined metric read-back
W_, H_ = i2c_b200.envs.QUAD_W, i2c_b200.envs.QUAD_H
Tm = 30
z_traj = np.zeros((Tm, 8))
z_traj[:, 0] = np.linspace(W_ / 4, 3 * W_ / 4, Tm)
z_traj[:, 1] = H_ / 2
Qm, Rm = np.diag([1e3, 1e3, 1e3, 1, 1, 1]), np.diag([1e-3, 1e-3])
u_init = 0.5 * 9.81 * i2c_b200.envs.QUAD_MASS * np.ones((6, 2))
gq = i2c_b200.BatchedI2c("Quadrotor", 48, 6, Qm, Rm, Qm / 1e3, 1.0, 1.0, u_init, 1e-2 * np.eye(2))
gq._propagate = True
pol = i2c_b200.BatchedPartiallyObservedMpc(gq, 2, 1e-2 * np.eye(2), z_traj, sig_zeta=np.diag([1e-6] * 8), pinned_io=True)
pol.set_control(feedforward=False)
gq.calibrate_alpha()
pol.optimize(3)
e = gq.env
y0 = np.array([e.x0[0] - 0.8, e.x0[1], e.x0[0] + 0.8, e.x0[1], 0, 0, 0.8, 0.8])
u = np.zeros((48, 2))
for t in range(4):
    u = pol(t, y0 + 1e-3 * rng.normal(size=(48, 8)), u)
print("mpc ok", bool(np.all(gq.status()[0] == 0)), flush=True)
gp = make_case(i2c_b200, "PendulumKnown", 96, 12)
out = [capi.pinned_empty((2, 96)), capi.pinned_empty((2, 96))]
for i in range(3):
    gp.run(1, capi.PH_LEARN, collect=False)
    gp.last_metrics_async(["alpha", "cost_m"], out[i & 1], i & 1)
    if i:
        gp.metrics_wait((i - 1) & 1)
gp.metrics_wait(0)
print("done", flush=True)
