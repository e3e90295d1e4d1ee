"""Parallel-in-time sweep (csrc/i2c_scan.cuh, i2c_run_scan): chunked associative-scan formulation of the Linearize
filter / smoother on linear systems.  Not in the reference (sequential loops, i2c/i2c.py:876-886); parity is against the
sequential kernel (itself pinned to the reference goldens in test_gpu_lqr.py), the NumPy oracle and the Riccati solution."""
import numpy as np
import pytest

from conftest import golden, relerr
from test_gpu_parity import i2c_b200  # noqa: F401
from test_gpu_lqr import finite_horizon_lqr, lqr_graph

pytestmark = pytest.mark.gpu

FIELDS = ["mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn", "mu_xu0_m", "sig_xu0_m", "K", "k", "sigK"]


def well_conditioned(m, B, T, seed, aux=False):
    rng = np.random.default_rng(seed)
    A = np.array([[1.0, 0.1], [-0.05, 0.98]]) + 0.01 * rng.normal(size=(B, 2, 2))
    Bm = np.array([[0.0], [0.1]])
    xg = rng.normal(size=(B, 2))
    a = xg - np.einsum("bij,bj->bi", A, xg)
    par = m.envs.linear_params(A, Bm, a)
    x0 = xg + 2.0 * rng.normal(size=(B, 2))
    z = np.repeat(np.concatenate((xg, np.zeros((B, 1))), axis=1)[:, None, :], T, axis=1)
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    kw = dict(x0=x0, sig_x0=1e-2 * np.eye(2), sig_eta=1e-3 * np.eye(2), env_par=par, z=z, z_term=xg, z_per_problem=True,
              inference="linearize", enable_aux=aux)
    return lambda: m.BatchedI2c("LinearKnown", B, T, np.diag([1.0, 2.0]), np.diag([0.5]), np.diag([1.0, 2.0]), 5.0, 0.5, mu_u,
                                np.eye(1), **kw)


@pytest.mark.parametrize("T,chunk", [(96, 16), (1000, 32), (257, 64)])
def test_scan_matches_sequential_em(i2c_b200, T, chunk):
    """EM iterations (alpha updates, prior <- posterior with feedback cells) through both kernels."""
    capi = i2c_b200.capi
    make = well_conditioned(i2c_b200, 40, T, 3)
    Gs, Gp = make(), make()
    for G in (Gs, Gp):
        G.set_cell_flag(capi.CELL_EXPERT, False)  # feedback cells without the pdf-ratio weighting stay linear-Gaussian
    Gp.time_parallel_chunk = chunk
    for it in range(4):
        Gs.learn(1)
        Gp.learn(1)
        assert np.all(Gs.status()[0] == 0) and np.all(Gp.status()[0] == 0), (it, Gp.status())
        for f in FIELDS:
            assert relerr(Gp.field(f), Gs.field(f)) < 1e-9, (it, f, relerr(Gp.field(f), Gs.field(f)))
        assert relerr(Gp.alpha, Gs.alpha) < 1e-11
    for name in ["alpha", "alpha_desired", "cost_m", "cost_m_var", "policy_entropy", "x_prior_entropy"]:
        assert relerr(np.array(Gp.metrics[name]), np.array(Gs.metrics[name])) < 1e-10, name
    # the records are interchangeable: continue the scan run with the sequential kernel
    Gp.time_parallel_chunk = None
    Gs.learn(1)
    Gp.learn(1)
    assert relerr(Gp.field("K"), Gs.field("K")) < 1e-9


def test_scan_refuses_what_it_cannot_do_exactly(i2c_b200):
    capi = i2c_b200.capi
    G = well_conditioned(i2c_b200, 8, 64, 0)()
    G.time_parallel_chunk = 16
    G.learn(1)  # all cells independent: fine; afterwards they are expert feedback cells (reference default)
    with pytest.raises(capi.I2cError, match="expert"):
        G.learn(1)
    rng = np.random.default_rng(0)
    P = i2c_b200.BatchedI2c("PendulumKnown", 8, 64, np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), np.diag([1.0, 100.0, 1.0]),
                            100.0, 0.0, 1e-2 * rng.normal(size=(8, 64, 1)), 2.0 * np.eye(1))
    P.time_parallel_chunk = 16
    with pytest.raises(capi.I2cError, match="linear"):
        P.learn(1)


def test_scan_lqr_vs_riccati_and_oracle(i2c_b200):
    """BASELINE config 2 through the parallel-in-time sweep: ill-conditioned by construction (sig_x0 = sig_eta = 1e-20)."""
    from oracle import envs as E
    from oracle import i2c_oracle as O

    g = golden("lqr_linearize")
    rng = np.random.default_rng(1)
    B, H = 64, int(g["H"])
    x0 = np.array([5.0, 5.0]) + rng.normal(size=(B, 2))
    A = g["A"] + 0.02 * rng.normal(size=(B, 2, 2))
    xag = np.broadcast_to(g["xag"], (B, 2)).copy()
    G, a = lqr_graph(i2c_b200, g, B, A, xag, x0)
    G.time_parallel_chunk = 12
    G.forward_backward(1)
    assert np.all(G.status()[0] == 0), G.status()
    K, k, _ = G.get_local_linear_policy()
    for b in range(0, B, 7):
        Kl, kl = finite_horizon_lqr(H, A[b], a[b], g["B"], g["Q"], g["R"], xag[b], np.zeros(1))
        assert np.max(np.abs(K[b] - Kl)) < 1e-5 * np.max(np.abs(Kl))
        assert np.max(np.abs(k[b] - kl)) < 1e-4 * np.max(np.abs(kl))
    R = O.Graph(E.Linear(A=A, B=g["B"], xg=xag), H, g["Q"], g["R"], g["Qf"], 1e-5, 0.0, np.zeros((H, 1)), 1e2 * np.eye(1), None,
                None, O.Linearize(), B=B, x0=x0)
    R._forward_backward_msgs()
    for name, tol in [("mu_xu1_f", 1e-8), ("sig_xu1_f", 1e-5), ("mu_xu0_m", 1e-7), ("sig_xu0_m", 1e-5), ("K", 1e-5), ("k", 1e-5)]:
        assert relerr(G.field(name), R.stack(name)) < tol, name
    G.time_parallel_chunk = None
    G.backward_ricatti()  # the Riccati sweep runs on the records the scan wrote
    lam = G.field("lambda_x3_b")[0] * 1e-5
    Gs, _ = lqr_graph(i2c_b200, g, B, A, xag, x0)
    Gs.forward_backward(1)
    Gs.backward_ricatti()
    assert relerr(lam, Gs.field("lambda_x3_b")[0] * 1e-5) < 1e-6


def test_scan_long_horizon_speedup(i2c_b200):
    """The point of the variant: a long horizon with few problems is latency-bound in the sequential kernel."""
    T = 4096
    make = well_conditioned(i2c_b200, 32, T, 5)
    Gs, Gp = make(), make()
    Gp.time_parallel_chunk = 64
    ms = {}
    for name, G in (("seq", Gs), ("scan", Gp)):
        G.forward_backward(2)
        G.synchronize()
        G.forward_backward(5)
        G.synchronize()
        ms[name] = G.last_run_ms() / 5
    assert relerr(Gp.field("K"), Gs.field("K")) < 1e-9
    print("ms per sweep pair:", ms)
    assert ms["scan"] < 0.5 * ms["seq"], ms


def test_scan_minimum_energy_covariance_control(i2c_b200):
    """LinearKnownMinimumEnergy (cost on u only: the state is never observed, J = 0 in every element) with covariance
    control at the end of the chain (linear_gaussian_covariance_control.py flow, no in-loop propagate) through both kernels."""
    capi = i2c_b200.capi
    g = golden("linear_covctrl_linearize")
    # (T = 60: this system is unstable and its state is never observed, so the filtered covariance grows like 1.1^(2T);
    #  at T = 200 it reaches 5e9 and BOTH kernels lose seven digits in the smoother differences -- they then agree to 1e-7)
    T = 60
    rng = np.random.default_rng(2)
    mu_u = 1e-2 * rng.normal(size=(T, 1))

    def make():
        G = i2c_b200.BatchedI2c("LinearKnownMinimumEnergy", 17, T, None, g["R"], None, float(g["alpha0"]), float(g["tol"]), mu_u,
                                g["sig_u"], g["mu_x_term"], g["sig_x_term"], inference="linearize", enable_aux=True)
        G.set_cell_flag(capi.CELL_EXPERT, False)
        return G

    Gs, Gp = make(), make()
    Gp.time_parallel_chunk = 16
    for it in range(3):
        Gs.learn(1)
        Gp.learn(1)
        assert np.all(Gp.status()[0] == 0), Gp.status()
        for f in FIELDS:
            assert relerr(Gp.field(f), Gs.field(f), floor=1e-9) < 1e-10, (it, f, relerr(Gp.field(f), Gs.field(f), floor=1e-9))
        assert relerr(Gp.alpha, Gs.alpha) < 1e-10
    assert Gp.temp == Gs.temp


def test_scan_argument_checks(i2c_b200):
    """Misuse is reported through the C-ABI error string, never silently redirected to another path."""
    capi = i2c_b200.capi
    G = well_conditioned(i2c_b200, 4, 64, 0)()
    for n_iter, phases, chunk, needle in [
        (1, capi.PH_FORWARD | capi.PH_BACKWARD, 4, "chunk_cells"),
        (1, capi.PH_FORWARD | capi.PH_BACKWARD, 65, "chunk_cells"),
        (1, capi.PH_FORWARD | capi.PH_BACKWARD | capi.PH_PROPAGATE, 16, "supports"),
        (1, capi.PH_BACKWARD | capi.PH_MSTEP, 16, "MSTEP needs FORWARD"),
        (0, capi.PH_FORWARD, 16, "n_iter"),
    ]:
        with pytest.raises(capi.I2cError, match=needle):
            capi.check(G.lib.i2c_run_scan(G._h, n_iter, phases, chunk))
    capi.check(G.lib.i2c_run_scan(G._h, 1, capi.PH_FORWARD | capi.PH_BACKWARD, 16))  # and the handle is still usable
    assert np.all(G.status()[0] == 0)
