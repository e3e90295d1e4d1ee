"""Linearize path on linear systems (BASELINE config 2): one forward/backward pass reproduces
finite-horizon LQR; golden = scripts/lqr_compare.py logic on the unmodified reference; batch = per-problem A, goal, x0."""
import numpy as np
import pytest

from conftest import golden, relerr
from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def finite_horizon_lqr(H, A, a, B, Q, R, xg, ug):
    """Ground truth: discrete-time finite-horizon LQR with affine term (restates i2c/utils.py:59-100)."""
    du, dx = B.shape[1], A.shape[0]
    K = np.zeros((H, du, dx))
    k = np.zeros((H, du))
    P, p = Q.copy(), -Q @ xg
    for i in range(H - 1, -1, -1):
        Minv = np.linalg.inv(R + B.T @ P @ B)
        K[i] = -Minv @ B.T @ P @ A
        k[i] = -Minv @ (B.T @ P @ a + B.T @ p - R @ ug)
        P_new = Q + A.T @ P @ A - A.T @ P @ B @ Minv @ B.T @ P @ A
        p = A.T @ (P @ a + p - P @ B @ Minv @ (B.T @ (P @ a + p) - R @ ug)) - Q @ xg
        P = P_new
    return K, k


def lqr_graph(m, g, B, A, xag, x0, aux=True):
    a = xag - np.einsum("bij,bj->bi", A, xag)
    par = m.envs.linear_params(A, g["B"], a)
    H = int(g["H"])
    z = np.concatenate((xag, np.zeros((B, 1))), axis=1)  # zg = [xag; 0]
    z_cells = np.repeat(z[:, None, :], H, axis=1)
    G = m.BatchedI2c("LinearKnown", B, H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1), x0=x0,
                     env_par=par, z=z_cells, z_term=xag, z_per_problem=True, inference="linearize", enable_aux=aux)
    return G, a


def riccati_noise_floor(g, n=6):
    """Round-off floor of the reference's Riccati messages on the LQR config: largest change of each field over n runs
    of the oracle with A perturbed by relative 2^-52 noise (a few ulps)."""
    from oracle import envs as E
    from oracle import i2c_oracle as O

    H = int(g["H"])
    rng = np.random.default_rng(0)

    def run(A):
        sys_ = E.Linear(A=A[None], B=g["B"], xg=g["xag"][None])
        G = O.Graph(sys_, H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1), None, None,
                    O.Linearize(), B=1, x0=g["x0"][None])
        for c in G.cells:
            c.state_action_independence = True
        G._forward_backward_msgs()
        G._backward_ricatti_msgs()
        return {a: G.stack(a)[0] for a in ("lambda_x3_b", "K", "k")}

    base = run(g["A"])
    floor = {a: 0.0 for a in base}
    for _ in range(n):
        out = run(g["A"] * (1.0 + 2.0 ** -52 * rng.integers(-2, 3, size=g["A"].shape)))
        for a in base:
            floor[a] = max(floor[a], float(np.max(np.abs(out[a] - base[a])) / np.max(np.abs(base[a]))))
    return floor


def test_lqr_golden_single(i2c_b200):
    g = golden("lqr_linearize")
    capi = i2c_b200.capi
    H = int(g["H"])
    e = i2c_b200.envs.make("LinearKnown")
    par = i2c_b200.envs.linear_params(g["A"], g["B"], g["a"])
    zg = np.concatenate((g["xag"], [0.0]))
    G = i2c_b200.BatchedI2c("LinearKnown", 1, H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1),
                            x0=g["x0"], env_par=par, z=np.repeat(zg[None], H, 0), z_term=g["xag"], inference="linearize",
                            enable_aux=True)
    G.z_graph[:] = zg
    G.forward_backward(1)  # lqr_compare.py:171 (all cells independent: the constructor state)
    assert np.all(G.status()[0] == 0), G.status()
    # sig_x0 = sig_eta = 1e-20 I, alpha = 1e-5: measured <= 5.8e-10 on every field
    for a, tol in [("mu_xu1_f", 1e-9), ("sig_xu1_f", 5e-9), ("mu_x3_f", 1e-9), ("sig_x3_f", 5e-9), ("mu_xu0_m", 1e-9),
                   ("sig_xu0_m", 5e-9), ("K", 5e-9), ("k", 5e-9), ("sigK", 1e-9), ("J_dyn", 5e-9)]:
        assert relerr(G.field(a)[0], g[f"fb/{a}"]) < tol, a
    K, k, _ = G.get_local_linear_policy()
    assert np.max(np.abs(K[0] - g["K_lqr"])) < 1e-5 * np.max(np.abs(g["K_lqr"]))
    assert np.max(np.abs(k[0] - g["k_lqr"])) < 1e-4 * np.max(np.abs(g["k_lqr"]))
    G.backward_ricatti()  # lqr_compare.py:175
    # The Riccati messages invert information matrices built from covariances of 1e-20 .. 1e2 (condition ~1e11): the
    # noise floor is what the REFERENCE arithmetic itself moves by when its input A is perturbed by a few ulps.
    # Tolerance = max(1e-9, 10 x that floor), measured on the oracle (same LAPACK calls as the reference).
    floor = riccati_noise_floor(g)
    assert relerr(G.field("lambda_x3_b")[0], g["ric/lambda_x3_b"]) < max(1e-9, 10 * floor["lambda_x3_b"])
    assert relerr(G.field("K")[0], g["ric/K"]) < max(1e-9, 10 * floor["K"]), floor
    assert relerr(G.field("k")[0], g["ric/k"]) < max(1e-9, 10 * floor["k"]), floor
    # value function: lambda_x3_b * alpha ~ P of the Riccati recursion (lqr_compare.py:85-110): a limit (measured 2.8e-8)
    P = g["P"]
    lam = G.field("lambda_x3_b")[0] * 1e-5
    assert relerr(lam, P) < 1e-6


def test_lqr_batched_8192_vs_riccati(i2c_b200):
    """BASELINE config 2: 8192 perturbed linear systems, each checked against its own Riccati solution."""
    g = golden("lqr_linearize")
    rng = np.random.default_rng(0)
    B = 8192
    x0 = np.array([5.0, 5.0]) + rng.normal(size=(B, 2))
    xag = np.array([10.0, 10.0]) + rng.normal(size=(B, 2))
    A = g["A"] + 0.02 * rng.normal(size=(B, 2, 2))
    G, a = lqr_graph(i2c_b200, g, B, A, xag, x0, aux=False)
    G.forward_backward(1)
    assert np.all(G.status()[0] == 0)
    K, k, _ = G.get_local_linear_policy()
    H = int(g["H"])
    idx = rng.choice(B, 512, replace=False)
    worst_K = worst_k = 0.0
    for b in idx:
        Kl, kl = finite_horizon_lqr(H, A[b], a[b], g["B"], g["Q"], g["R"], xag[b], np.zeros(1))
        worst_K = max(worst_K, np.max(np.abs(K[b] - Kl)) / np.max(np.abs(Kl)))
        worst_k = max(worst_k, np.max(np.abs(k[b] - kl)) / np.max(np.abs(kl)))
    assert worst_K < 1e-5 and worst_k < 1e-4, (worst_K, worst_k)


def test_lqr_batched_vs_oracle(i2c_b200):
    from oracle import i2c_oracle as O
    from oracle import envs as E

    g = golden("lqr_linearize")
    rng = np.random.default_rng(1)
    B, H = 48, int(g["H"])
    x0 = np.array([5.0, 5.0]) + rng.normal(size=(B, 2))
    A = g["A"] + 0.02 * rng.normal(size=(B, 2, 2))
    xag = np.broadcast_to(g["xag"], (B, 2)).copy()
    G, a = lqr_graph(i2c_b200, g, B, A, xag, x0)
    sys_ = E.Linear(A=A, B=g["B"], xg=xag)
    R = O.Graph(sys_, H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1), None, None, O.Linearize(),
                B=B, x0=x0)
    G.forward_backward(1)
    R._forward_backward_msgs()
    # 48 perturbed systems (some nearly unstable): measured <= 2.2e-9 over all fields
    for name, tol in [("mu_xu1_f", 1e-9), ("sig_xu1_f", 2e-8), ("mu_xu0_m", 1e-9), ("sig_xu0_m", 2e-8), ("K", 2e-8),
                      ("k", 2e-8), ("mu_z0_m", 1e-9), ("sig_z0_m", 2e-8), ("mu_x3_m", 1e-9)]:
        assert relerr(G.field(name), R.stack(name)) < tol, name


def test_linear_covariance_control_linearize(i2c_b200):
    """linear_gaussian_covariance_control.py:91-125 flow (Linearize + propagate + covariance control)."""
    capi = i2c_b200.capi
    g = golden("linear_covctrl_linearize")
    T = int(g["T"])
    G = i2c_b200.BatchedI2c("LinearKnownMinimumEnergy", 1, T, None, g["R"], None, float(g["alpha0"]), float(g["tol"]),
                            g["mu_u"], g["sig_u"], g["mu_x_term"], g["sig_x_term"], inference="linearize", enable_aux=True)
    G.set_cell_flag(capi.CELL_EXPERT, False)
    G._propagate = True
    for it in range(1, 6):
        G.learn(1)
        assert np.all(G.status()[0] == 0), G.status()
        if it <= 2:
            for a in ["mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK", "mu_x3_pf",
                      "sig_x3_pf"]:
                assert relerr(G.field(a)[0], g[f"it{it}/{a}"], floor=1e-9) < 1e-9, (it, a)  # measured <= 2e-11
    assert relerr(np.array(G.metrics["cost_m"])[:, 0], g["costs_m"]) < 1e-9
    assert relerr(np.array(G.metrics["kl_term"])[:, 0], g["kl_terms"]) < 1e-9


# gains of the Linearize path on the nonlinear envs: max(1e-9, ~10 x measured); measured 1.5e-12 / 1.8e-10 / 7.4e-10 against
# both the oracle and the unmodified reference (states, covariances: <= 4e-13)
TOL_LIN_GAIN = {"PendulumKnown": 1e-9, "CartpoleKnown": 2e-9, "DoubleCartpoleKnown": 8e-9}


@pytest.mark.parametrize("env,Q,R,alpha,xs,T", [
    ("PendulumKnown", np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), 100.0, [0.3, 0.5], 40),
    ("CartpoleKnown", np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0]), 80.0, 0.05, 40),
    ("DoubleCartpoleKnown", 1e-3 * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0]), 1e-4 * np.eye(1), 0.05, 0.02, 30),
])
def test_linearize_nonlinear_envs(i2c_b200, env, Q, R, alpha, xs, T):
    """Linearize inference on the nonlinear envs (experiments pendulum_known.py,
    cartpole_known.py, double_cartpole_known(_lin).py).  The kernel takes the dynamics Jacobians by forward-mode AD
    (csrc/dual.cuh); the reference by autograd, the oracle -- and the reference run that produced the golden of problem 0
    (tests/golden/*_linearize_*.npz, unmodified reference code with a complex-step autograd.jacobian) -- by the
    complex-step derivative: all three are the exact derivative to rounding."""
    from oracle import i2c_oracle as O

    rng = np.random.default_rng(4)
    e = i2c_b200.envs.make(env)
    B = 16
    x0 = e.x0 + np.asarray(xs) * rng.normal(size=(B, e.dim_x))
    mu_u = 1e-2 * rng.normal(size=(B, T, e.dim_u))
    G = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, alpha, 0.5, mu_u, np.eye(e.dim_u), x0=x0, inference="linearize",
                            enable_aux=True)
    ref = O.make_graph(env, T, Q, R, Q, alpha, 0.5, mu_u, np.eye(e.dim_u), inference=O.Linearize(), B=B, x0=x0)
    gname = {"PendulumKnown": "pendulum_linearize_T40", "CartpoleKnown": "cartpole_linearize_T40",
             "DoubleCartpoleKnown": "double_cartpole_linearize_T30"}[env]
    g = golden(gname)
    assert np.array_equal(g["x0"], x0[0]) and np.array_equal(g["mu_u"], mu_u[0])
    tol_g = TOL_LIN_GAIN[env]
    for it in range(3):
        G.learn(1)
        ref.learn_msgs()
        assert np.all(G.status()[0] == 0), (it, G.status())
        for a in ["mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "mu_xu0_m", "sig_xu0_m", "mu_z0_m", "sig_z0_m", "sigK"]:
            assert relerr(G.field(a), ref.stack(a)) < 1e-9, (it, a)
            assert relerr(G.field(a)[0], g[f"it{it + 1}/{a}"]) < 1e-9, (it, a)  # the unmodified reference
        for a in ["J_dyn", "K", "k"]:
            assert relerr(G.field(a), ref.stack(a), floor=1e-6) < tol_g, (it, a)
            assert relerr(G.field(a)[0], g[f"it{it + 1}/{a}"], floor=1e-6) < tol_g, (it, a)
        assert relerr(G.alpha, ref.alpha) < 1e-9
    assert relerr(np.array(G.metrics["cost_m"]), np.array(ref.costs_m)) < 1e-9
    assert relerr(np.array(G.metrics["cost_m"])[:, 0], g["costs_m"]) < 1e-9
    assert relerr(np.array(G.metrics["alpha"])[:, 0], g["alphas"][1:]) < 1e-9
