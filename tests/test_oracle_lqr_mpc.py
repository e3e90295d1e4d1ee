"""Oracle pinned to the reference's Linearize/LQR path (scripts/lqr_compare.py logic) and to the
partially observed MPC loop (policy/mpc.py) -- goldens from tests/golden/make_golden.py."""
import numpy as np
import pytest

from conftest import golden, relerr
from oracle import i2c_oracle as O
from oracle import envs as E


def lqr_graph(g, B=1, A=None, xag=None, x0=None):
    A = g["A"] if A is None else A
    xag = g["xag"] if xag is None else xag
    sys_ = E.Linear(A=A, B=g["B"], xg=xag)
    H = int(g["H"])
    G = O.Graph(sys_, H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1), None, None,
                O.Linearize(), B=B, x0=g["x0"] if x0 is None else x0)
    for c in G.cells:
        c.state_action_independence = True
    return G


def test_lqr_linearize_pass():
    g = golden("lqr_linearize")
    G = lqr_graph(g)
    G._forward_backward_msgs()
    # conditioning: sig_x0 = sig_eta = 1e-20 I, alpha = 1e-5 (SURVEY.md section 7 "LQR config conditioning")
    for a, tol in [("mu_xu1_f", 1e-9), ("sig_xu1_f", 1e-7), ("mu_x3_f", 1e-9), ("sig_x3_f", 1e-7), ("mu_xu0_m", 1e-9),
                   ("sig_xu0_m", 1e-7), ("K", 1e-7), ("k", 1e-7), ("sigK", 1e-9)]:
        assert relerr(G.stack(a)[0], g[f"fb/{a}"]) < tol, a
    K, k, _ = G.get_local_linear_policy()
    # LQR equivalence is a limit (alpha -> 0, sig_u -> inf): 2.6e-7 / 3.8e-6 measured on the reference
    assert np.max(np.abs(K[0] - g["K_lqr"])) < 1e-5 * np.max(np.abs(g["K_lqr"]))
    assert np.max(np.abs(k[0] - g["k_lqr"])) < 1e-4 * np.max(np.abs(g["k_lqr"]))
    G._backward_ricatti_msgs()
    for a, tol in [("lambda_x3_b", 1e-5), ("K", 1e-5), ("k", 1e-5)]:
        assert relerr(G.stack(a)[0], g[f"ric/{a}"]) < tol, a


def test_linear_covariance_control_linearize():
    g = golden("linear_covctrl_linearize")
    sys_ = E.LinearMinimumEnergy()
    T = int(g["T"])
    G = O.Graph(sys_, T, None, g["R"], None, float(g["alpha0"]), float(g["tol"]), g["mu_u"], g["sig_u"],
                g["mu_x_term"], g["sig_x_term"], O.Linearize())
    for c in G.cells:
        c.use_expert_controller = False
    G._propagate = True
    for it in range(1, 6):
        G.learn_msgs()
        if it <= 2:
            for a in ["mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK",
                      "mu_x3_pf", "sig_x3_pf"]:
                assert relerr(G.stack(a)[0], g[f"it{it}/{a}"], floor=1e-9) < 1e-8, (it, a)
    assert relerr(np.array([a[0] for a in G.costs_m]), g["costs_m"]) < 1e-8
    assert relerr(np.array([a[0] for a in G.kl_terms]), g["kl_terms"]) < 1e-7


@pytest.mark.parametrize("mode", ["ff_low", "ff_high", "fb_low", "fb_high"])
def test_mpc_quadrotor(mode):
    g = golden(f"mpc_quadrotor_{mode}")
    sys_ = E.Quadrotor(sig_zeta=g["sig_zeta"])
    T_plan = int(g["T_plan"])
    G = O.Graph(sys_, T_plan, g["Q"], g["R"], g["Qf"], 1.0, 1.0, g["u_init"], g["sig_u"], None, None, O.Cubature(1, 0, 0))
    G._propagate = True
    pol = O.PartiallyObservedMpc(G, int(g["mpc_iter"]), g["sig_u"], g["z_traj"].copy())
    pol.set_control(bool(g["feedforward"]))
    G.calibrate_alpha()
    assert abs(G.alpha[0] - g["alpha_cal1"]) < 1e-10 * g["alpha_cal1"]
    pol.optimize(25, G.x0, G.sig_x0)
    G.calibrate_alpha()
    assert abs(G.alpha[0] - g["alpha_cal2"]) < 1e-8 * g["alpha_cal2"]
    assert relerr(G.stack("mu_xu0_m")[0], g["warm/mu_xu0_m"]) < 1e-9
    assert relerr(G.stack("sig_xu0_m")[0], g["warm/sig_xu0_m"]) < 1e-8
    u = np.zeros((1, 2))
    x = sys_.x0.copy()
    n_steps = g["u"].shape[0]
    for t in range(n_steps):
        y = g["y"][t][None]
        u = pol(t, y, u)
        assert relerr(pol.mu[0], g["mu"][t]) < 1e-8, t
        assert relerr(pol.covar[0], g["covar"][t]) < 1e-7, t
        u = np.clip(u, 0.0, 30.0)
        assert relerr(u[0], g["u"][t]) < 1e-7, t
