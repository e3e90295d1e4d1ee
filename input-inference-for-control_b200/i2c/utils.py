"""Host-side helpers the scripts import next to the hot path (subset mirror of i2c/utils.py): the finite-horizon
LQR ground truth used by the LQR-equivalence experiment, seeding and logging.  Plot / results-folder helpers of the
reference are out of scope."""
import logging
import random

import numpy as np


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)


def setup_logger(res_dir=None, level=logging.INFO):
    logging.basicConfig(level=level)
    return logging.getLogger()


def finite_horizon_lqr(H, A, a, B, Q, R, x0, xg, ug, dim_x, dim_u):
    """Affine finite-horizon LQR by the backward Riccati recursion (interface of i2c/utils.py:59-100):
    returns x_lqr, u_lqr, K, k, cost, Ps, ps."""
    K = np.zeros((H, dim_u, dim_x))
    k = np.zeros((H, dim_u))
    Ps = np.zeros((H, dim_x, dim_x))
    ps = np.zeros((H, dim_x))
    P, p = np.array(Q, float), -Q @ xg
    for i in reversed(range(H)):
        Ps[i], ps[i] = P, p
        Minv = np.linalg.inv(R + B.T @ P @ B)
        K[i] = -Minv @ B.T @ P @ A
        k[i] = -Minv @ (B.T @ P @ a + B.T @ p - R @ ug)
        Pa_p = P @ a + p
        p = A.T @ (Pa_p - P @ B @ Minv @ (B.T @ Pa_p - R @ ug)) - Q @ xg
        P = Q + A.T @ P @ A - A.T @ P @ B @ Minv @ B.T @ P @ A
    x_lqr, u_lqr = np.zeros((H, dim_x)), np.zeros((H, dim_u))
    x, cost = np.array(x0, float), 0.0
    for i in range(H):
        x_lqr[i] = x
        u = K[i] @ x + k[i]
        u_lqr[i] = u
        cost += (x - xg) @ Q @ (x - xg) + (u - ug) @ R @ (u - ug)
        x = A @ x + B @ u + a
    cost += (x - xg) @ Q @ (x - xg)
    return x_lqr, u_lqr, K, k, cost, Ps, ps
