"""Gauss-Hermite quadrature (exp_types.py:52-68, quadrature.py:132).  Goldens come from the unmodified
reference (tests/golden/make_golden.py gh): gauss_hermite_kat.npz, pendulum_gh3_T40.npz, pendulum_gh4_propagate_T20.npz.
CPU tests pin the oracle and the library's node rule; GPU tests compare the grid kernels with both."""
import numpy as np
import pytest

from conftest import GAINS, golden, relerr

KAT = [("PendulumKnown", 4), ("PendulumKnown", 3), ("LinearKnown", 5), ("CartpoleKnown", 3)]


# ------------------------------------------------------------------------------------------- CPU: oracle + host rule
def test_oracle_gh_rule_matches_reference():
    from oracle import i2c_oracle as O

    g = golden("gauss_hermite_kat")
    for deg in (1, 2, 3, 4, 7):
        sf, w, _ = O.GaussHermite(deg).weights(2)
        assert np.array_equal(np.concatenate(([sf], w)), g[f"weights/{deg}"])
        assert np.array_equal(O.GaussHermite(deg).pts(2), g[f"pts/{deg}"])


@pytest.mark.parametrize("env,deg", KAT)
def test_oracle_gh_transforms(env, deg):
    from oracle import envs as E
    from oracle import i2c_oracle as O

    g = golden("gauss_hermite_kat")
    sys_ = E.make(env)
    n, dx = sys_.dim_xu, sys_.dim_x
    t = f"{env}/{deg}"
    m, S = g[f"{t}/m_in"][None], g[f"{t}/S_in"][None]
    q = O.Quad(O.GaussHermite(deg), n)
    mz, Sz = q.forward(sys_.observe, m, S)
    assert relerr(mz[0], g[f"{t}/obs_m"]) < 1e-14 and relerr(Sz[0], g[f"{t}/obs_S"]) < 1e-11
    assert relerr(q.sig_xy[0], g[f"{t}/obs_Sxy"]) < 1e-11
    mx, Sx, Sn = q.forward_gaussian(sys_.forward, m, S)
    assert relerr(mx[0], g[f"{t}/dyn_m"]) < 1e-14 and relerr(Sx[0], g[f"{t}/dyn_S"]) < 1e-11
    assert relerr(q.sig_xy[0], g[f"{t}/dyn_Sxy"]) < 1e-11 and relerr(Sn, g[f"{t}/dyn_Sn"]) < 1e-14


@pytest.mark.parametrize("name,deg", [("pendulum_gh3_T40", 3), ("pendulum_gh4_propagate_T20", 4)])
def test_oracle_gh_em_against_reference(name, deg):
    from oracle import i2c_oracle as O
    from test_oracle_golden import opt, run_em_case

    g = golden(name)
    G = O.make_graph(str(g["env"]), int(g["T"]), opt(g["Q"]), g["R"], opt(g["Qf"]), float(g["alpha0"]), float(g["tol"]),
                     g["mu_u"], g["sig_u"], opt(g["mu_x_term"]), opt(g["sig_x_term"]), inference=O.GaussHermite(deg), B=1,
                     x0=g["x0"])
    run_em_case(G, g, 1e-10, 5e-9)


def test_library_gauss_hermite_rule():
    """The C library's own node rule (no GPU needed) against numpy's hermgauss, which the reference calls."""
    import __graft_entry__ as ge

    ge.build()
    import i2c_b200

    for deg in range(1, 9):
        x, w = i2c_b200.gauss_hermite(deg)
        xr, wr = np.polynomial.hermite.hermgauss(deg)
        assert np.max(np.abs(x - xr)) < 5e-16 * max(1.0, np.max(np.abs(xr)))
        assert np.max(np.abs(w * np.sqrt(np.pi) - wr)) < 1e-15
        assert np.array_equal(x, -x[::-1]) and np.array_equal(w, w[::-1])
    with pytest.raises(i2c_b200.I2cError):
        i2c_b200.gauss_hermite(9)


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def i2c_b200():
    import __graft_entry__ as ge

    ge.build()
    import i2c_b200 as m
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("env,deg", KAT)
def test_gpu_gh_transforms_vs_reference(i2c_b200, env, deg):
    g = golden("gauss_hermite_kat")
    t = f"{env}/{deg}"
    m, S = g[f"{t}/m_in"][None], g[f"{t}/S_in"][None]
    dx = {"PendulumKnown": 2, "LinearKnown": 2, "CartpoleKnown": 4}[env]
    my, Sy, Sxy, st = i2c_b200.quadrature(env, "observe", m, S, gh_degree=deg)
    assert st[0] == 0
    assert relerr(my[0], g[f"{t}/obs_m"]) < 1e-13 and relerr(Sy[0], g[f"{t}/obs_S"]) < 1e-11
    assert relerr(Sxy[0], g[f"{t}/obs_Sxy"]) < 1e-11
    my, Sy, Sxy, st = i2c_b200.quadrature(env, "forward", m, S, gh_degree=deg)
    assert relerr(my[0], g[f"{t}/dyn_m"]) < 1e-13 and relerr(Sy[0], g[f"{t}/dyn_S"]) < 1e-10
    assert relerr(Sxy[0], g[f"{t}/dyn_Sxy"]) < 1e-10
    my, Sy, Sxy, st = i2c_b200.quadrature(env, "observe_terminal", m[:, :dx], S[:, :dx, :dx], gh_degree=deg)
    assert relerr(my[0], g[f"{t}/term_m"]) < 1e-13 and relerr(Sy[0], g[f"{t}/term_S"]) < 1e-11


FIELDS = ["mu_z0_f", "sig_z0_f", "mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn", "mu_x3_m", "sig_x3_m", "mu_xu0_m",
          "sig_xu0_m", "mu_z0_m", "sig_z0_m", "K", "k", "sigK"]
PF = ["mu_xu0_pf", "sig_xu0_pf", "mu_z0_pf", "sig_z0_pf", "mu_x3_pf", "sig_x3_pf"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,deg", [("pendulum_gh3_T40", 3), ("pendulum_gh4_propagate_T20", 4)])
def test_gpu_gh_em_against_reference_golden(i2c_b200, name, deg):
    capi = i2c_b200.capi
    g = golden(name)
    G = i2c_b200.BatchedI2c(str(g["env"]), 1, int(g["T"]), g["Q"], g["R"], g["Qf"], float(g["alpha0"]), float(g["tol"]),
                            g["mu_u"], g["sig_u"], x0=g["x0"], enable_aux=True, inference="gauss_hermite", quadrature=deg)
    propagate = bool(g["propagate"])
    if propagate:
        G._propagate = True
        G.set_cell_flag(capi.CELL_EXPERT, bool(g["expert"]))
        G.run(1, capi.PH_PROPAGATE, False)
        for a in PF:
            assert relerr(G.field(a)[0], g[f"it0/{a}"]) < 1e-9, a
    n_dump, n_total = int(g["n_dump"]), int(g["n_total"])
    for it in range(1, n_total + 1):
        G.learn(1)
        assert np.all(G.status()[0] == 0), G.status()
        if it <= n_dump:
            for a in FIELDS + (PF if propagate else []):
                e = relerr(G.field(a)[0], g[f"it{it}/{a}"], floor=1e-6 if a in GAINS else 0.0)
                assert e < (1e-7 if a in GAINS else 1e-9), (it, a, e)
    assert relerr(np.array([a[0] for a in G.alphas]), g["alphas"]) < 1e-8
    assert relerr(np.array(G.metrics["cost_m"])[:, 0], g["costs_m"]) < 1e-8
    K, k, sk = G.get_local_linear_policy()
    assert relerr(K[0], g["final/K"], 1e-6) < 1e-6 and relerr(k[0], g["final/k"], 1e-6) < 1e-6
    if propagate:
        assert relerr(np.array(G.metrics["cost_pf"])[:, 0], g["costs_pf"]) < 1e-8
        assert relerr(np.array(G.metrics["alpha_pf"])[:, 0], g["alphas_pf"][1:]) < 1e-8


@pytest.mark.gpu
def test_gpu_gh_batched_vs_oracle_cartpole(i2c_b200):
    """Degree 3 on the cart-pole: 3^5 = 243 points per joint transform, 3^4 = 81 for the terminal one."""
    from oracle import i2c_oracle as O

    rng = np.random.default_rng(21)
    B, T = 33, 12
    e = i2c_b200.envs.make("CartpoleKnown")
    Q, R = np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0])
    x0 = e.x0 + 0.05 * rng.normal(size=(B, 4))
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    G = i2c_b200.BatchedI2c("CartpoleKnown", B, T, Q, R, Q, 80.0, 0.0, mu_u, np.eye(1), x0=x0, enable_aux=True,
                            inference="gauss_hermite", quadrature=3)
    ref = O.make_graph("CartpoleKnown", T, Q, R, Q, 80.0, 0.0, mu_u, np.eye(1), inference=O.GaussHermite(3), B=B, x0=x0)
    for it in range(2):
        G.learn(1)
        ref.learn_msgs()
        assert np.all(G.status()[0] == 0)
        for a in FIELDS:
            err = relerr(G.field(a), ref.stack(a), floor=1e-6 if a in GAINS else 0.0)
            assert err < (1e-6 if a in GAINS else 1e-9), (it, a, err)
    assert relerr(G.alpha, ref.alpha) < 1e-10


@pytest.mark.gpu
def test_gpu_gh_mirror(i2c_b200):
    """The drop-in surface: QuadratureInference(GaussHermiteQuadrature(4), dim) (quadrature.py:132) and an I2cGraph built
    with GaussHermiteQuadrature inference (i2c.py:116, 839)."""
    from i2c.exp_types import GaussHermiteQuadrature
    from i2c.i2c import I2cGraph
    from i2c.inference.quadrature import QuadratureInference
    from i2c.model import make_env_model

    g = golden("gauss_hermite_kat")
    sys_ = make_env_model("PendulumKnown", None)
    t = "PendulumKnown/4"
    q = QuadratureInference(GaussHermiteQuadrature(4), 3)
    assert q.n_points == 64 and q.base_pts.shape == (64, 3)
    m, S = q.forward(sys_.observe, g[f"{t}/m_in"][:, None], g[f"{t}/S_in"])
    assert relerr(m[:, 0], g[f"{t}/obs_m"]) < 1e-13 and relerr(S, g[f"{t}/obs_S"]) < 1e-11
    assert relerr(q.sig_xy, g[f"{t}/obs_Sxy"]) < 1e-11
    gg = golden("pendulum_gh3_T40")
    sys_.x0 = gg["x0"].reshape(-1, 1)
    graph = I2cGraph(sys_, int(gg["T"]), gg["Q"], gg["R"], gg["Qf"], float(gg["alpha0"]), float(gg["tol"]), gg["mu_u"],
                     gg["sig_u"], None, None, GaussHermiteQuadrature(3))
    for _ in range(int(gg["n_total"])):
        graph.learn_msgs()
    assert relerr(np.array(graph.alphas), gg["alphas"]) < 1e-8
    K, k, sk = graph.get_local_linear_policy()
    assert relerr(K, gg["final/K"], 1e-6) < 1e-6


@pytest.mark.gpu
def test_gpu_gh_argument_checks(i2c_b200):
    capi = i2c_b200.capi
    m, S = np.zeros((1, 3)), np.eye(3)[None]
    with pytest.raises(capi.I2cError, match="degree"):
        i2c_b200.quadrature("PendulumKnown", "observe", m, S, gh_degree=9)
    with pytest.raises(capi.I2cError, match="degree"):
        i2c_b200.BatchedI2c("PendulumKnown", 1, 8, np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), np.diag([1.0, 100.0, 1.0]), 100.0, 0.0,
                            np.zeros((8, 1)), 2.0 * np.eye(1), inference="gauss_hermite", quadrature=12)
