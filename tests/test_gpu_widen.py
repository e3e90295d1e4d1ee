"""GPU parity for the rows of SURVEY.md 8(a) beyond the plain EM sweep: closed-loop propagate + calibrate_alpha
(a12), covariance control, MPC with the cubature Kalman filter and horizon shift (a15, a16), plus the
boundary's state handling (status words, snapshot / restore, fused phases)."""
import numpy as np
import pytest

from conftest import GAINS, golden, relerr
from test_gpu_parity import ALIAS, FIELDS_B, FIELDS_F, FIELDS_P, compare_cells, i2c_b200, make_pair  # noqa: F401

pytestmark = pytest.mark.gpu


def opt(a):
    return None if a.size == 0 else a


def graph_from_golden(m, g, B=1, **kw):
    return m.BatchedI2c(str(g["env"]), B, int(g["T"]), opt(g["Q"]), g["R"], opt(g["Qf"]), float(g["alpha0"]),
                        float(g["tol"]), g["mu_u"], g["sig_u"], opt(g["mu_x_term"]), opt(g["sig_x_term"]), x0=g["x0"],
                        enable_aux=True, **kw)


# tolerances = max(1e-9, ~10 x measured).  pendulum_actreg_covctrl: cost on u only, alpha = 300, T = 100 with covariance
# control -- the reference's own restatement (oracle, same LAPACK) only reproduces this golden to 2e-9 (tests/
# test_oracle_golden.py); measured here: states 7.7e-9, gains 3.6e-9.  double cart-pole: 3e-10 / 4.5e-9.  pendulum expert: 2e-10.
@pytest.mark.parametrize("name,tol_s,tol_g", [("pendulum_actreg_covctrl_T100", 5e-8, 5e-8),
                                              ("double_cartpole_covctrl_T50", 3e-9, 5e-8),
                                              ("pendulum_propagate_expert_T50", 1e-9, 3e-9)])
def test_propagate_and_covariance_control_golden(i2c_b200, name, tol_s, tol_g):
    """nonlinear_covariance_control.py:81-115 flow against the unmodified reference (B = 1)."""
    capi = i2c_b200.capi
    g = golden(name)
    G = graph_from_golden(i2c_b200, g)
    G._propagate = True
    G.set_cell_flag(capi.CELL_EXPERT, bool(g["expert"]))
    G.propagate()
    for a in FIELDS_P:
        assert relerr(G.field(a)[0], g[f"it0/{a}"]) < tol_s, a
    n_dump, n_total = int(g["n_dump"]), int(g["n_total"])
    for it in range(1, n_dump + 1):
        G.learn(1)
        assert np.all(G.status()[0] == 0), G.status()
        for a in FIELDS_F + FIELDS_B + FIELDS_P:
            e = relerr(G.field(ALIAS.get(a, a))[0], g[f"it{it}/{a}"], floor=1e-6 if a in GAINS else 0.0)
            assert e < (tol_g if a in GAINS else tol_s), (it, a, e)
    G.learn(n_total - n_dump)
    m = G.metrics
    assert relerr(np.array(G.alphas)[:, 0], g["alphas"]) < tol_s
    for mine, ref in [("alpha_desired", "alphas_desired"), ("alpha_pf", "alphas_pf"), ("cost_m", "costs_m"),
                      ("cost_m_var", "costs_m_var"), ("cost_pf", "costs_pf"), ("cost_pf_var", "costs_pf_var"),
                      ("cost_pf_min", "cost_pf_min"), ("policy_entropy", "policy_entropy"),
                      ("x_prior_entropy", "x_prior_entropy"), ("propagate_entropy", "propagate_entropy")]:
        r = g[ref][1:] if ref.startswith("alphas") else g[ref]
        assert relerr(np.array(m[mine])[:, 0], r) < 2 * tol_s, mine
    if "kl_terms" in g.files:
        assert relerr(np.array(m["kl_term"])[:, 0], g["kl_terms"]) < 2 * tol_s
    K, k, sk = G.get_local_linear_policy()
    assert relerr(K[0], g["final/K"], 1e-6) < tol_g and relerr(k[0], g["final/k"], 1e-6) < tol_g


def test_covariance_control_batched_vs_oracle(i2c_b200):
    capi = i2c_b200.capi
    sf = 1e-3
    Q = sf * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0])
    mu_t, sig_t = np.zeros(6), np.diag([0.01, 0.005, 0.005, 0.05, 0.05, 0.05])
    G, ref = make_pair(i2c_b200, "DoubleCartpoleKnown", 24, 30, Q, sf * np.diag([0.1]), Q, 0.05, 0.99, 9, 0.02, np.eye(1),
                       mu_t, sig_t)
    G._propagate = ref._propagate = True
    G.set_cell_flag(capi.CELL_EXPERT, False)
    for c in ref.cells:
        c.use_expert_controller = False
    G.propagate()
    ref.propagate()
    for it in range(3):
        G.learn(1)
        ref.learn_msgs()
        assert np.all(G.status()[0] == 0)
        compare_cells(G, ref, FIELDS_F + FIELDS_B + FIELDS_P, 1e-8, 1e-5, tag=f"it{it}")
    assert G.temp == ref.cells[-1].temp
    assert relerr(np.array(G.metrics["kl_term"]), np.array(ref.kl_terms)) < 1e-6
    assert relerr(np.array(G.metrics["cost_pf"]), np.array(ref.costs_pf)) < 1e-8


def test_calibrate_alpha_and_split_phases(i2c_b200):
    """calibrate_alpha (i2c.py:895-911) and running the sweeps as separate calls (forward / backward /
    update_priors) equals the fused learn_msgs launch."""
    capi = i2c_b200.capi
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    G, ref = make_pair(i2c_b200, "PendulumKnown", 40, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1))
    G._propagate = ref._propagate = True
    G.calibrate_alpha()
    ref.calibrate_alpha()
    assert relerr(G.alpha, ref.alpha) < 1e-12
    for _ in range(2):
        G.learn(1)
        ref.learn_msgs()
    G.calibrate_alpha(only_decrease=True)
    ref.calibrate_alpha(only_decrease=True)
    assert relerr(G.alpha, ref.alpha) < 1e-10
    assert relerr(np.array(G.alphas), np.array(ref.alphas)) < 1e-10
    # split phases
    G2, _ = make_pair(i2c_b200, "PendulumKnown", 40, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1))
    G3, _ = make_pair(i2c_b200, "PendulumKnown", 40, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1))
    for _ in range(3):
        G2.forward_backward(1, update_priors=True)
        G3.forward()
        G3.backward()
        G3.update_priors()
    for a in ["mu_xu0_m", "sig_xu0_m", "K", "k", "sigK", "prior_mu", "prior_K"]:
        assert np.array_equal(G2.field(a), G3.field(a)), a
    assert np.array_equal(G2.get_cell_flags(), G3.get_cell_flags())
    assert not np.any(G2.get_cell_flags() & capi.CELL_INDEPENDENT)


def test_status_words_isolate_failures(i2c_b200):
    """A problem whose start covariance is not PD is flagged; its neighbours are untouched (the reference raises
    LinAlgError for the whole run, quadrature.py:17-24)."""
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    rng = np.random.default_rng(0)
    B, T = 70, 20
    x0 = np.array([np.pi, 0.0]) + 0.1 * rng.normal(size=(B, 2))
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    sig_x0 = np.broadcast_to(1e-5 * np.eye(2), (B, 2, 2)).copy()
    sig_x0[13] = -sig_x0[13]
    G = i2c_b200.BatchedI2c("PendulumKnown", B, T, Q, R, Q, 100.0, 0.0, mu_u, 2.0 * np.eye(1), x0=x0, sig_x0=sig_x0)
    G.learn(2)
    st, info = G.status()
    assert st[13] == 1 and np.all(np.delete(st, 13) == 0)
    Gok = i2c_b200.BatchedI2c("PendulumKnown", B, T, Q, R, Q, 100.0, 0.0, mu_u, 2.0 * np.eye(1), x0=x0)
    Gok.learn(2)
    keep = np.delete(np.arange(B), 13)
    assert np.array_equal(G.field("K")[keep], Gok.field("K")[keep])


def test_snapshot_restore(i2c_b200):
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    G, _ = make_pair(i2c_b200, "PendulumKnown", 33, 25, Q, R, Q, 100.0, 0.0, 2, np.array([0.3, 0.5]), 2.0 * np.eye(1))
    G.learn(2)
    snap = G.snapshot()
    G.learn(3)
    K5 = G.field("K")
    a5 = G.alpha
    G.restore(snap)
    G.learn(3)
    assert np.array_equal(G.field("K"), K5) and np.array_equal(G.alpha, a5)


def quad_setup(m, g, B, sig_zeta, pinned_io=False):
    G = m.BatchedI2c("Quadrotor", B, int(g["T_plan"]), g["Q"], g["R"], g["Qf"], 1.0, 1.0, g["u_init"], g["sig_u"],
                     enable_aux=True)
    G._propagate = True
    pol = m.BatchedPartiallyObservedMpc(G, int(g["mpc_iter"]), g["sig_u"], g["z_traj"].copy(), sig_zeta=sig_zeta,
                                        pinned_io=pinned_io)
    pol.set_control(bool(g["feedforward"]))
    return G, pol


@pytest.mark.parametrize("mode", ["ff_low", "ff_high", "fb_low", "fb_high"])
def test_mpc_quadrotor_golden(i2c_b200, mode):
    """mpc_quad.py:538-652 driver flow (calibrate, 25 warm-start sweeps, calibrate, closed loop) against the
    reference's PartiallyObservedMpcPolicy run with the same fp64 quadrotor restatement (B = 1)."""
    g = golden(f"mpc_quadrotor_{mode}")
    G, pol = quad_setup(i2c_b200, g, 1, g["sig_zeta"])
    G.calibrate_alpha()
    assert abs(G.alpha[0] - g["alpha_cal1"]) < 1e-10 * g["alpha_cal1"]
    pol.optimize(25)
    G.calibrate_alpha()
    assert abs(G.alpha[0] - g["alpha_cal2"]) < 1e-8 * g["alpha_cal2"]
    assert relerr(G.field("mu_xu0_m")[0], g["warm/mu_xu0_m"]) < 1e-8
    assert relerr(G.field("sig_xu0_m")[0], g["warm/sig_xu0_m"]) < 1e-7
    u = np.zeros((1, 2))
    for t in range(g["u"].shape[0]):
        u = pol(t, g["y"][t][None], u)
        mu, cov = pol.belief
        assert relerr(mu[0], g["mu"][t]) < 1e-7, t
        assert relerr(cov[0], g["covar"][t]) < 1e-6, t
        u = np.clip(u, 0.0, 30.0)
        assert relerr(u[0], g["u"][t]) < 1e-6, t
    assert np.all(G.status()[0] == 0)


def test_mpc_batched_rollouts_vs_oracle(i2c_b200):
    """B independent closed-loop roll-outs with per-roll-out noise: CUDA path vs batched oracle."""
    from oracle import i2c_oracle as O
    from oracle import envs as E

    g = golden("mpc_quadrotor_fb_high")
    B, n_steps = 12, 6
    sig_zeta = g["sig_zeta"]
    G, pol = quad_setup(i2c_b200, g, B, sig_zeta)
    sys_ = E.Quadrotor(sig_zeta=sig_zeta)
    R = O.Graph(sys_, int(g["T_plan"]), g["Q"], g["R"], g["Qf"], 1.0, 1.0, g["u_init"], g["sig_u"], None, None,
                O.Cubature(1, 0, 0), B=B)
    R._propagate = True
    rp = O.PartiallyObservedMpc(R, int(g["mpc_iter"]), g["sig_u"], g["z_traj"].copy())
    rp.set_control(bool(g["feedforward"]))
    for obj, p in ((G, pol), (R, rp)):
        obj.calibrate_alpha()
        if obj is G:
            p.optimize(25)
        else:
            p.optimize(25, R.x0, R.sig_x0)
        obj.calibrate_alpha()
    assert relerr(G.alpha, R.alpha) < 1e-8
    rng = np.random.default_rng(3)
    x = np.broadcast_to(sys_.x0, (B, 6)).copy()
    u = np.zeros((B, 2))
    ur = np.zeros((B, 2))
    G2, pol2 = quad_setup(i2c_b200, g, B, sig_zeta)  # same roll-outs through the un-fused call sequence
    G2.calibrate_alpha()
    pol2.optimize(25)
    G2.calibrate_alpha()
    u2 = np.zeros((B, 2))
    # ... and through the fused call with page-locked I/O (i2c_host_alloc): measurements from a pinned array, the returned
    # actions (views into the policy's pinned ring) fed back as they are
    G3, pol3 = quad_setup(i2c_b200, g, B, sig_zeta, pinned_io=True)
    G3.calibrate_alpha()
    pol3.optimize(25)
    G3.calibrate_alpha()
    y3 = i2c_b200.capi.pinned_empty((B, 8))
    u3 = np.zeros((B, 2))
    for t in range(n_steps):
        y = sys_.measure(x) + rng.multivariate_normal(np.zeros(8), sig_zeta, B)
        u2 = np.clip(pol2(t, y, u2, fused=False), 0.0, 30.0)
        u = np.clip(pol(t, y, u), 0.0, 30.0)
        assert np.array_equal(u, u2), t
        y3[:] = y
        u3 = pol3(t, y3, u3)
        np.clip(u3, 0.0, 30.0, out=u3)
        assert np.array_equal(u, u3), t
        ur = np.clip(rp(t, y, ur), 0.0, 30.0)
        assert relerr(u, ur) < 1e-6, t
        mu, cov = pol.belief
        assert relerr(mu, rp.mu) < 1e-7 and relerr(cov, rp.covar) < 1e-6, t
        x = sys_.dynamics(np.concatenate((x, ur), axis=-1)) + rng.multivariate_normal(np.zeros(6), sys_.sig_eta, B)
        u = u2 = ur
        u3[:] = ur
    assert np.all(G.status()[0] == 0)
    assert np.array_equal(G.get_cell_flags(), G3.get_cell_flags()) and np.array_equal(G.get_cell_flags(), G2.get_cell_flags())


def test_async_policy_copy(i2c_b200):
    """i2c_get_policy_async + i2c_copy_wait deliver the same controllers as the synchronous getter, also when a new
    sweep is launched while the copy is in flight."""
    import torch

    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    G, _ = make_pair(i2c_b200, "PendulumKnown", 300, 40, Q, R, Q, 100.0, 0.0, 7, np.array([0.3, 0.5]), 2.0 * np.eye(1),
                     enable_aux=False)
    G.learn(2)
    K0, k0, s0 = G.get_local_linear_policy()
    pin = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    K, k, s = pin(300, 40, 1, 2), pin(300, 40, 1), pin(300, 40, 1, 1)
    G.get_local_linear_policy_async(K, k, s)
    G.learn(1)  # overlaps the copy; must not disturb it
    G.wait_copies()
    assert np.array_equal(K, K0) and np.array_equal(k, k0) and np.array_equal(s, s0)
    K1, k1, s1 = G.get_local_linear_policy()
    G.get_local_linear_policy_async(K, k, s)
    G.wait_copies()
    assert np.array_equal(K, K1) and not np.array_equal(K1, K0)


def test_initial_state_async_upload(i2c_b200):
    """i2c_set_initial_state_async (no trailing synchronisation; the caller keeps the pinned buffers alive) gives the same
    sweep as the synchronous call."""
    import torch

    capi = i2c_b200.capi
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    Ga, _ = make_pair(i2c_b200, "PendulumKnown", 70, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1), enable_aux=False)
    Gb, _ = make_pair(i2c_b200, "PendulumKnown", 70, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1), enable_aux=False)
    rng = np.random.default_rng(9)
    x0 = torch.empty((70, 2), dtype=torch.float64, pin_memory=True).numpy()
    s0 = torch.empty((70, 2, 2), dtype=torch.float64, pin_memory=True).numpy()
    x0[:] = np.array([np.pi, 0.0]) + 0.2 * rng.normal(size=(70, 2))
    s0[:] = 1e-4 * np.eye(2)
    capi.check(Ga.lib.i2c_set_initial_state(Ga._h, capi.ptr(x0), capi.ptr(s0)))
    capi.check(Gb.lib.i2c_set_initial_state_async(Gb._h, capi.ptr(x0), capi.ptr(s0)))
    Ga.learn(2)
    Gb.learn(2)
    assert np.array_equal(Ga.field("K"), Gb.field("K")) and np.array_equal(Ga.alpha, Gb.alpha)
    xa, sa = Ga.get_initial_state()
    assert np.array_equal(xa, x0) and relerr(sa, s0) < 1e-15
    xb, sb = Gb.get_initial_state()
    assert np.array_equal(xb, xa) and np.array_equal(sb, sa)
    # the asynchronous call fills alternate belief buffers on an upload stream (overlapping a running sweep) and swaps them in:
    # a stream of new beliefs queued back to back with the sweeps, a snapshot / restore and a deepcopy in the swapped state
    x1 = [capi.pinned_empty((70, 2)) for _ in range(3)]
    for i, x in enumerate(x1):
        x[:] = x0 + 0.01 * (i + 1)
    for x in x1:
        capi.check(Ga.lib.i2c_set_initial_state(Ga._h, capi.ptr(x), capi.ptr(s0)))
        Ga.learn(1)
        capi.check(Gb.lib.i2c_set_initial_state_async(Gb._h, capi.ptr(x), capi.ptr(s0)))
        Gb.run(1, capi.PH_LEARN, collect=False)  # queued behind the upload, no host synchronisation in between
    assert np.array_equal(Ga.field("K"), Gb.field("K")) and np.array_equal(Ga.alpha, Gb.alpha)
    assert np.array_equal(Gb.get_initial_state()[0], x1[-1])
    snap = Gb.snapshot()
    Gb.learn(1)
    K1 = Gb.field("K")
    capi.check(Gb.lib.i2c_set_initial_state_async(Gb._h, capi.ptr(x1[0]), capi.ptr(s0)))  # swapped state again
    Gb.restore(snap)
    assert np.array_equal(Gb.get_initial_state()[0], x1[-1])
    Gb.learn(1)
    Ga.learn(1)
    assert np.array_equal(Gb.field("K"), K1) and np.array_equal(Ga.field("K"), K1)


def test_pipelined_metrics_read(i2c_b200):
    """i2c_get_last_metrics_async / i2c_metrics_wait (two staging slots, copy stream): a loop that queues step i+1 before it
    collects step i sees exactly the numbers of the synchronous getter."""
    capi = i2c_b200.capi
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    Ga, _ = make_pair(i2c_b200, "PendulumKnown", 500, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1), enable_aux=False)
    Gb, _ = make_pair(i2c_b200, "PendulumKnown", 500, 30, Q, R, Q, 100.0, 0.0, 5, np.array([0.3, 0.5]), 2.0 * np.eye(1), enable_aux=False)
    names = ["alpha", "cost_m", "policy_entropy"]
    out = [capi.pinned_empty((3, 500)), capi.pinned_empty((3, 500))]
    got = []
    n = 7
    for i in range(n):
        Ga.run(1 if i % 3 else 2, capi.PH_LEARN, collect=False)  # (the async getter reads the LAST iteration of the run)
        Ga.last_metrics_async(names, out[i & 1], i & 1)
        if i > 0:
            Ga.metrics_wait((i - 1) & 1)
            got.append(out[(i - 1) & 1].copy())
    Ga.metrics_wait((n - 1) & 1)
    got.append(out[(n - 1) & 1].copy())
    for i in range(n):
        Gb.run(1 if i % 3 else 2, capi.PH_LEARN, collect=True)
        ref = np.stack([np.array(Gb.metrics[m][-1]) for m in names])
        assert np.array_equal(got[i], ref), i
