"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's policy-evaluation roll-out
(BaseSim.run, i2c/env.py:40-74; BaseKnownSim.forward, :180-187) under the time-indexed linear-Gaussian policies
(i2c/policy/linear.py:31-43, 73-90), batched over problems and roll-outs.  Disturbances are passed in explicitly
(the reference draws them from the global NumPy RNG)."""
import numpy as np


def rollout(sys, x_init, K, k, eta, sig_k=None, eps_u=None, expert=None, soft=True, hard_threshold=3.0):
    """x_init [B,R,dx], K [B,T,du,dx], k [B,T,du], eta [B,R,T,dx]; returns xu [B,R,T,n], z [B,R,T,dz], z_term [B,R,dzt]."""
    B, R, dx = x_init.shape
    T, du = K.shape[1], K.shape[2]
    x = x_init.copy()
    xu_t, z_t = [], []
    for t in range(T):
        if expert is not None:
            mu, lam = expert
            d = x - mu[:, None, t, :]
            e = 0.5 * np.einsum("bri,bij,brj->br", d, lam[:, t], d)
            gate = np.exp(-e) if soft else (np.abs(e) < hard_threshold).astype(float)
            u = k[:, None, t, :] + gate[..., None] * np.einsum("bij,brj->bri", K[:, t], d)
        else:
            u = np.einsum("bij,brj->bri", K[:, t], x) + k[:, None, t, :]
        if sig_k is not None:
            Ls = np.linalg.cholesky(sig_k[:, t])
            u = u + np.einsum("bij,brj->bri", Ls, eps_u[:, :, t, :])
        xu = np.concatenate((x, u), axis=-1)
        xu_t.append(xu)
        z_t.append(sys.observe(xu))
        x = sys.dynamics(xu) + eta[:, :, t, :]
    zt = sys.observe_terminal(x)
    return np.stack(xu_t, axis=2), np.stack(z_t, axis=2), zt
