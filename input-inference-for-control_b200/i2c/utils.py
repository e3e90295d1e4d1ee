"""Host-side helpers the scripts import next to the hot path (subset mirror of i2c/utils.py): the finite-horizon
LQR ground truth used by the LQR-equivalence experiment, seeding and logging.  Plot / results-folder helpers of the
reference are out of scope."""
import logging
import random

import numpy as np


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)


DATETIME = __import__("datetime").datetime.now().strftime("%Y-%m-%d-%H-%M-%S")


def make_results_folder(config, seed, name, folder_name="_results", release=False):
    """`_results/<release|timestamp>_<config>_<seed>_<name>` (i2c/utils.py:355-366)."""
    import os

    parts = [config.replace(" ", "-"), str(seed), name.replace(" ", "-")]
    res_dir = os.path.join(folder_name, "_".join((["release"] if release else [DATETIME]) + parts))
    os.makedirs(res_dir, exist_ok=True)
    return res_dir


def configure_plots():
    """Plot styling of the reference (i2c/utils.py:369-377); a no-op without matplotlib (plotting is outside the path)."""
    try:
        import matplotlib

        matplotlib.rcParams["font.family"] = "serif"
        matplotlib.rcParams["figure.figsize"] = [16, 16]
        matplotlib.rcParams["legend.fontsize"] = 16
        matplotlib.rcParams["axes.titlesize"] = 22
        matplotlib.rcParams["axes.labelsize"] = 22
    except Exception:
        pass


def covariance_2d(covar, mean, axis, n_std=2.0, facecolor="b", **kwargs):
    """n_std ellipse of a 2-D Gaussian added to a matplotlib axis (i2c/utils.py:380-397)."""
    w, v = np.linalg.eig(np.asarray(covar, float))
    width = 2 * n_std * np.sqrt(w)
    assert not np.any(np.isnan(width))
    theta = np.rad2deg(-np.arctan2(v[0, 1], v[0, 0]))
    try:
        from matplotlib.patches import Ellipse

        return axis.add_patch(Ellipse(xy=np.asarray(mean).squeeze(), width=width[0], height=width[1], angle=theta,
                                      edgecolor=facecolor, facecolor="none", **kwargs))
    except Exception:
        return None


def write_commit(res_dir):
    """git_commit.txt with the current revision (i2c/utils.py:423-430); skipped outside a git checkout."""
    import os
    import subprocess

    try:
        sha = subprocess.check_output(["git", "rev-parse", "HEAD"], stderr=subprocess.DEVNULL, text=True).strip()
    except Exception:
        sha = "unknown"
    with open(os.path.join(res_dir, "git_commit.txt"), "w+") as f:
        f.write(sha)


def setup_logger(res_dir=None, level=logging.INFO):
    """i2c/utils.py:400-408: log into <res_dir>/output.log (console when no folder is given)."""
    import os

    for handler in logging.root.handlers[:]:
        logging.root.removeHandler(handler)
    if res_dir:
        logging.basicConfig(filename=os.path.join(res_dir, "output.log"), level=level,
                            format="[%(asctime)s] %(pathname)s:%(lineno)d %(levelname)s- %(message)s")
    else:
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


def _quadratic_traj_cost(W, Wf, zg, zg_term, dim_x, z, z_term):
    """sum_{t < T-1} (z_t - zg)^T W (z_t - zg) [+ terminal (x_T - xg)^T Wf (x_T - xg)]: the last step of the trajectory is
    dropped and only the first dim_x terminal features count, as in i2c/utils.py:116-121, 160-168."""
    e = np.asarray(z, float).reshape(-1, W.shape[0])[:-1] - zg.reshape(1, -1)
    cost = float(np.sum((e @ W) * e))
    if z_term is not None:
        et = (np.asarray(z_term, float).reshape(1, -1) - zg_term.reshape(1, -1))[-1, :dim_x]
        cost += float(et @ Wf @ et)
    return cost


class TrajectoryEvaluator(object):
    """Cost of one executed roll-out next to the planned one (interface of i2c/utils.py:103-148)."""

    def __init__(self, W, Wf, sg, sg_term, dim_x):
        self.W, self.Wf = np.asarray(W, float), np.asarray(Wf, float)
        self.sg, self.sg_term = np.asarray(sg, float).reshape(-1, 1), np.asarray(sg_term, float).reshape(-1, 1)
        self.dim_x = dim_x
        self.actual_cost, self.planned_cost = [], []
        if self.W.shape != (self.sg.shape[0],) * 2:
            raise AssertionError("W must be square and match the feature goal")

    def _eval_traj(self, s, s_terminal):
        return _quadratic_traj_cost(self.W, self.Wf, self.sg, self.sg_term, self.dim_x, s, np.asarray(s_terminal)[-1])

    def eval(self, actual_traj, actual_terminal, planned_traj, planned_terminal):
        self.actual_cost.append(self._eval_traj(actual_traj, actual_terminal))
        self.planned_cost.append(self._eval_traj(planned_traj, planned_terminal))

    def plot(self, name, res_dir=None):
        logging.info("TrajectoryEvaluator.plot: plotting is outside the CUDA path (SURVEY.md section 2)")

    def save(self, name, res_dir):
        import os

        np.save(os.path.join(res_dir, f"cost_actual_{name}.npy"), np.asarray(self.actual_cost))
        np.save(os.path.join(res_dir, f"cost_plan_{name}.npy"), np.asarray(self.planned_cost))


class StochasticTrajectoryEvaluator(object):
    """Cost statistics (mean, min, max, 10th / 90th percentile) over a batch of roll-outs, e.g. the output of
    ``env.batch_eval`` (interface of i2c/utils.py:151-265; file names of ``save`` as read by process_results.py)."""

    def __init__(self, W, Wf, sg, sg_term, dim_x):
        self.W, self.Wf = np.asarray(W, float), np.asarray(Wf, float)
        self.sg, self.sg_term = np.asarray(sg, float).reshape(-1, 1), np.asarray(sg_term, float).reshape(-1, 1)
        self.dim_x, self.dim_s = dim_x, self.W.shape[0]
        self.mu_actual_cost, self.max_actual_cost, self.min_actual_cost = [], [], []
        self.actual_cost_10, self.actual_cost_90, self.planned_cost = [], [], []
        if self.W.shape[1] != self.dim_s or self.sg.shape[0] != self.dim_s:
            raise AssertionError(f"{self.sg.shape[0]} ~= {self.dim_s}")

    def _eval_traj(self, s, s_term):
        return _quadratic_traj_cost(self.W, self.Wf, self.sg, self.sg_term, self.dim_x, s, s_term)

    def eval(self, actual_trajs, actual_trajs_term, planned_traj, planned_traj_term):
        costs = np.array([self._eval_traj(z, zt) for z, zt in zip(actual_trajs, actual_trajs_term)])
        self.mu_actual_cost.append(costs.mean())
        self.min_actual_cost.append(costs.min())
        self.max_actual_cost.append(costs.max())
        lo, hi = np.percentile(costs, (10, 90))
        self.actual_cost_10.append(lo)
        self.actual_cost_90.append(hi)
        self.planned_cost.append(self._eval_traj(planned_traj, planned_traj_term))

    def plot(self, name, res_dir=None):
        logging.info("StochasticTrajectoryEvaluator.plot: plotting is outside the CUDA path (SURVEY.md section 2)")

    plot_sample = plot

    def save(self, name, res_dir):
        import os

        np.save(os.path.join(res_dir, f"cost_actual_mean_{name}.npy"), np.asarray(self.mu_actual_cost))
        np.save(os.path.join(res_dir, f"cost_plan_{name}.npy"), np.asarray(self.planned_cost))
