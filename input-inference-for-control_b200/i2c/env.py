"""Evaluation simulators (mirror of the reference's i2c/env.py interface: make_env, run, batch_eval) on the GPU:
all roll-outs of a policy run in ONE launch of the batched roll-out kernel (i2c_rollout) instead of a Python loop
per step and a multiprocessing pool per batch (i2c/env.py:40-103).  Disturbances are drawn on the host from the
global NumPy RNG (as the reference does) so that seeding behaves the same; rendering / plotting are out of scope."""
import numpy as np

import i2c_b200
from i2c.policy.linear import ExpertTimeIndexedLinearGaussianPolicy


class KnownSim(object):
    simulated = True
    deterministic = False  # process noise is always on (env_def.py BaseDef.deterministic)
    env = None

    def __init__(self, env_name, duration):
        c = i2c_b200.envs.make(env_name)
        self._name = env_name
        self.duration = duration
        self.dim_x, self.dim_u, self.dim_z, self.dim_z_term = c.dim_x, c.dim_u, c.dim_z, c.dim_z_term
        self.dim_s = self.dim_xu = c.dim_x + c.dim_u
        self.x0 = c.x0.reshape(-1, 1).copy()
        self.sig_x0 = c.sig_x0.copy()
        self.sig_eta = c.sig_eta.copy()
        self._random_start = env_name.startswith("Linear")  # BaseLinear.init_env draws x ~ N(x0, sig_x0)
        self._par = i2c_b200.envs.linear_params(c.A, c.B, c.a) if hasattr(c, "A") else None

    def _rollouts(self, policy, n, deterministic):
        T = self.duration
        K, k = policy.K[None, :T], policy.k[None, :T]
        sig_k = None if deterministic else policy.sig_k[None, :T]
        expert = None
        soft = True
        if isinstance(policy, ExpertTimeIndexedLinearGaussianPolicy):
            expert, soft = (policy.mu[None, :T], policy.lam[None, :T]), policy.soft
        if self._random_start:
            x_init = np.random.multivariate_normal(self.x0[:, 0], self.sig_x0, n)[None]
        else:
            x_init = np.broadcast_to(self.x0[:, 0], (1, n, self.dim_x)).copy()
        eta = np.random.multivariate_normal(np.zeros(self.dim_x), self.sig_eta, (1, n, T))
        eps_u = None if deterministic else np.random.standard_normal((1, n, T, self.dim_u))
        xu, z, zt, xf = i2c_b200.rollout(self._name, x_init, K, k, sig_k=sig_k, expert=expert, soft_expert=soft, eta=eta,
                                         eps_u=eps_u, env_par=self._par, return_final=True)
        x_next = np.concatenate((xu[0, :, 1:, : self.dim_x], xf[0, :, None, :]), axis=1)
        y = x_next - xu[0, :, :, : self.dim_x]
        return xu[0], y, z[0], zt[0]

    def run(self, policy, deterministic=True, render=False, use_tqdm=False):
        """One roll-out: returns (xt [T, dim_s], yt [T, dim_x], zt [T, dim_z], z_term [1, dim_z_term])."""
        xu, y, z, zt = self._rollouts(policy, 1, deterministic)
        return xu[0], y[0], z[0], zt[0][None, :]

    def run_render(self, policy, dir, name="", deterministic=True, use_tqdm=False):
        return self.run(policy, deterministic)

    def batch_eval(self, policy, n_eval, deterministic=True):
        xu, y, z, zt = self._rollouts(policy, n_eval, deterministic)
        return ([xu[i] for i in range(n_eval)], [y[i] for i in range(n_eval)], [z[i] for i in range(n_eval)],
                [zt[i][None, :] for i in range(n_eval)])

    def plot_sim(self, *args, **kwargs):
        pass

    def plot_trajectory(self, *args, **kwargs):
        pass

    def close(self):
        pass


def make_env(exp):
    """Simulator for an experiment module / object with ENVIRONMENT and N_DURATION (i2c/env.py:17-32)."""
    name = {"PendulumKnownLearn": "PendulumKnown", "CartpoleSetKnown": "CartpoleKnown", "CartpoleKnownLearn": "CartpoleKnown",
            "DoubleCartpoleKnownLearn": "DoubleCartpoleKnown"}.get(exp.ENVIRONMENT, exp.ENVIRONMENT)
    return KnownSim(name, exp.N_DURATION)
