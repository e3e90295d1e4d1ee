"""The drop-in mirror (`from i2c.i2c import I2cGraph` ...) driven the way the reference's scripts drive the reference
(scripts/i2c_run.py:29-131, lqr_compare.py:120-176, nonlinear_covariance_control.py:81-115,
mpc_state_est/mpc_quad.py:538-652), checked against goldens produced by the unmodified reference."""
import copy
import pickle

import numpy as np
import pytest

from conftest import GAINS, golden, relerr
from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mirror(i2c_b200):
    from i2c.exp_types import CubatureQuadrature, GaussianI2c, Linearize
    from i2c.i2c import I2cGraph
    from i2c.inference.quadrature import QuadratureInference
    from i2c.model import make_env_model, QuadrotorKnown
    from i2c.policy.linear import ExpertTimeIndexedLinearGaussianPolicy, TimeIndexedLinearGaussianPolicy
    from i2c.policy.mpc import PartiallyObservedMpcPolicy
    from i2c.utils import finite_horizon_lqr
    import types

    return types.SimpleNamespace(**locals())


def test_i2c_run_inference_loop(mirror):
    """scripts/i2c_run.py:29-47,89-98 with experiments/pendulum_known_quad.py (seed 0)."""
    g = golden("pendulum_known_quad_seed0")
    model = mirror.make_env_model("PendulumKnown", None)
    inf = mirror.GaussianI2c(inference=mirror.CubatureQuadrature(1, 0, 0), Q=g["Q"], R=g["R"], Qf=g["Qf"], alpha=100,
                             alpha_update_tol=0.0, mu_u=g["mu_u"], sig_u=g["sig_u"], mu_x_term=None, sig_x_term=None)
    i2c = mirror.I2cGraph(model, 100, inf.Q, inf.R, inf.Qf, inf.alpha, inf.alpha_update_tol, inf.mu_u, inf.sig_u,
                          inf.mu_x_term, inf.sig_x_term, inf.inference, res_dir=None)
    policy_linear = mirror.TimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), 100, i2c.sys.dim_u, i2c.sys.dim_x)
    policy = mirror.ExpertTimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), 100, i2c.sys.dim_u, i2c.sys.dim_x, soft=False)
    i2c.reset_metrics()
    for i in range(200):
        i2c.learn_msgs()
        if i < 3:
            c = i2c.cells[10]
            for a in ["mu_x0_f", "sig_x0_f", "mu_xu1_f", "sig_xu1_f", "mu_x3_f", "sig_x3_f", "J_dyn", "mu_x3_m", "sig_x3_m",
                      "mu_xu0_m", "sig_xu0_m", "mu_z0_m", "sig_z0_m", "K", "k", "sigK", "mu_xu0_f"]:
                mine = np.asarray(getattr(c, a))
                ref = g[f"it{i + 1}/{a}"][10]
                assert relerr(mine.reshape(ref.shape), ref, 1e-6 if a in GAINS else 0.0) < (1e-7 if a in GAINS else 1e-9), a
        policy_linear.write(*i2c.get_local_linear_policy())
        policy.write(*i2c.get_local_expert_linear_policy())
    assert i2c.alphas[0] == 100 and len(i2c.alphas) == 201 and len(i2c.costs_m) == 200
    # golden schedule quoted in SURVEY.md section 6
    assert abs(i2c.alphas[1] - 99.23063222598958) < 1e-9 and abs(i2c.alphas[2] - 93.70997137355181) < 1e-8
    assert relerr(np.array(i2c.alphas), g["alphas"]) < 1e-8
    assert relerr(np.array(i2c.costs_m), g["costs_m"]) < 1e-8
    assert abs(i2c.costs_m[0] - 39606.560956855086) < 1e-6
    assert relerr(policy_linear.K, g["final/K"]) < 1e-6
    assert relerr(policy_linear.k, g["final/k"]) < 1e-6
    u = policy_linear(0, np.asarray(model.x0, float))
    assert u.shape == (1, 1)
    z_est, z_term_est = i2c.get_marginal_observed_trajectory()
    assert z_est.shape == (100, 4) and z_term_est.shape == (3, 1)
    assert i2c.get_marginal_trajectory().shape == (100, 3)
    # deepcopy / pickle round trip keeps the device state (policy/mpc.py:24, i2c.py:1392-1401)
    clone = copy.deepcopy(i2c)
    blob = pickle.dumps(i2c)
    i2c.learn_msgs()
    clone.learn_msgs()
    again = pickle.loads(blob)
    again.learn_msgs()
    for other in (clone, again):
        assert other.alphas[-1] == i2c.alphas[-1]
        assert np.array_equal(other.get_local_linear_policy()[0], i2c.get_local_linear_policy()[0])


def test_lqr_compare_flow(mirror):
    """scripts/lqr_compare.py:120-176."""
    g = golden("lqr_linearize")
    model = mirror.make_env_model("LinearKnown", None)
    model.xag = 10 * np.ones((2, 1))
    model.zg_term = 10 * np.ones((2, 1))
    model.a = model.xag - model.A @ model.xag
    H = int(g["H"])
    x_lqr, u_lqr, K_lqr, k_lqr, cost_lqr, P, p = mirror.finite_horizon_lqr(
        H, model.A, model.a[:, 0], model.B, g["Q"], g["R"], model.x0[:, 0], model.xag[:, 0], np.zeros((1,)), 2, 1)
    assert relerr(K_lqr, g["K_lqr"]) < 1e-14 and relerr(k_lqr, g["k_lqr"]) < 1e-13 and relerr(P, g["P"]) < 1e-14
    i2c = mirror.I2cGraph(sys=model, horizon=H, Q=g["Q"], R=g["R"], Qf=g["Qf"], alpha=1e-5, alpha_update_tol=0.0,
                          mu_u=np.zeros((H, 1)), sig_u=1e2 * np.eye(1), mu_x_terminal=None, sig_x_terminal=None,
                          inference=mirror.Linearize(), res_dir=None)
    i2c.use_expert_controller = False
    for c in i2c.cells:
        c.state_action_independence = True
    i2c._forward_backward_msgs()
    K, k, _ = i2c.get_local_linear_policy()
    assert np.max(np.abs(K - K_lqr)) < 1e-5 * np.max(np.abs(K_lqr))
    assert np.max(np.abs(k - k_lqr)) < 1e-4 * np.max(np.abs(k_lqr))
    x, u = i2c.get_state_and_action()
    assert np.max(np.abs(x[:, :, 0] - x_lqr)) < 1e-5 and np.max(np.abs(u[:, :, 0] - u_lqr)) < 1e-4
    i2c._backward_ricatti_msgs()
    lam = np.asarray([c.lambda_x3_b for c in i2c.cells]) * i2c.alpha
    assert relerr(lam, P) < 1e-3


def test_nonlinear_covariance_control_flow(mirror):
    """scripts/nonlinear_covariance_control.py:81-115 with experiments/pendulum_known_act_reg_quad.py."""
    g = golden("pendulum_actreg_covctrl_T100")
    model = mirror.make_env_model("PendulumKnownActReg", None)
    i2c = mirror.I2cGraph(sys=model, horizon=100, Q=None, R=g["R"], Qf=None, alpha=300.0, alpha_update_tol=1.0,
                          mu_u=g["mu_u"], sig_u=g["sig_u"], mu_x_terminal=g["mu_x_term"], sig_x_terminal=g["sig_x_term"],
                          inference=mirror.CubatureQuadrature(1, 0, 0), res_dir=None)
    for c in i2c.cells:
        c.use_expert_controller = False
    i2c._propagate = True
    i2c.propagate()
    for i in range(15):
        i2c.learn_msgs()
    c = i2c.cells[-1]
    assert c.mu_x3_m.shape == (2, 1) and c.sig_x3_pf.shape == (2, 2)
    assert relerr(np.array(i2c.kl_terms), g["kl_terms"]) < 3e-8  # measured 2.2e-9 (reference's own floor 2e-9, see test_gpu_widen)
    assert relerr(np.array(i2c.costs_pf), g["costs_pf"]) < 1e-8
    assert relerr(np.array(i2c.alphas), g["alphas"]) < 1e-12  # tol = 1.0: alpha frozen, exact
    K, k, s = i2c.get_local_linear_policy()
    assert relerr(K, g["final/K"], 1e-6) < 2e-8  # measured 1.4e-9


def test_quadrature_inference_mirror(mirror):
    g = golden("quadrature_kat")
    model = mirror.make_env_model("CartpoleKnown", None)
    q = mirror.QuadratureInference(mirror.CubatureQuadrature(1, 0, 0), model.dim_xu)
    m, S = q.forward(model.observe, g["CartpoleKnown/0/m_in"][:, None], g["CartpoleKnown/0/S_in"])
    assert m.shape == (6, 1) and relerr(m[:, 0], g["CartpoleKnown/0/obs_m"]) < 1e-13
    assert relerr(S, g["CartpoleKnown/0/obs_S"]) < 1e-11 and relerr(q.sig_xy, g["CartpoleKnown/0/obs_Sxy"]) < 1e-11
    m, S, Sn = q.forward_gaussian(model.forward, g["CartpoleKnown/0/m_in"][:, None], g["CartpoleKnown/0/S_in"])
    assert relerr(m[:, 0], g["CartpoleKnown/0/dyn_m"]) < 1e-13 and relerr(Sn, g["CartpoleKnown/0/dyn_Sn"]) < 1e-15
    with pytest.raises(NotImplementedError):  # arbitrary callables cannot run in-kernel: no CPU fallback
        q.forward(lambda x: x, np.zeros((5, 1)), np.eye(5))
    with pytest.raises(np.linalg.LinAlgError):
        q.forward(model.observe, np.zeros((5, 1)), -np.eye(5))


@pytest.mark.parametrize("mode", ["ff_low", "fb_high"])
def test_mpc_quad_flow(mirror, mode):
    """scripts/mpc_state_est/mpc_quad.py:538-652 (i2c branch), closed loop with the golden's noise draws."""
    g = golden(f"mpc_quadrotor_{mode}")
    model = mirror.QuadrotorKnown()
    model.sig_zeta = g["sig_zeta"]
    T_plan = int(g["T_plan"])
    _i2c = mirror.I2cGraph(sys=model, horizon=T_plan, Q=g["Q"], R=g["R"], Qf=g["Qf"], alpha=1.0, alpha_update_tol=1.0,
                           mu_u=g["u_init"], sig_u=g["sig_u"], mu_x_terminal=None, sig_x_terminal=None,
                           inference=mirror.CubatureQuadrature(1, 0, 0), res_dir=None)
    _i2c._propagate = True
    policy = mirror.PartiallyObservedMpcPolicy(_i2c, int(g["mpc_iter"]), g["sig_u"], np.copy(g["z_traj"]))
    policy.set_control(feedforward=bool(g["feedforward"]))
    policy.i2c.calibrate_alpha()
    assert abs(policy.i2c.alpha - g["alpha_cal1"]) < 1e-10 * g["alpha_cal1"]
    policy.optimize(25, model.x0, model.sig_x0)
    policy.i2c.calibrate_alpha()
    assert abs(policy.i2c.alpha - g["alpha_cal2"]) < 1e-8 * g["alpha_cal2"]
    u = np.zeros((model.dim_u, 1))
    for t in range(g["u"].shape[0]):
        u = policy(t, g["y"][t][:, None], u)
        u = model.clip_u(u.T).T
        assert relerr(u[:, 0], g["u"][t]) < 1e-6, t
        assert relerr(policy.mus[-1][:, 0], g["mu"][t]) < 1e-7, t
        assert relerr(policy.xu_history[-1][:, :, 0], g["plan"][t]) < 1e-6, t


def test_env_batch_eval_mirror(mirror):
    """scripts/i2c_run.py:96-106: evaluate the extracted controllers in the simulator after every EM iteration."""
    import types

    from i2c.env import make_env

    g = golden("pendulum_known_quad_seed0")
    exp = types.SimpleNamespace(ENVIRONMENT="PendulumKnown", N_DURATION=100)
    env = make_env(exp)
    model = mirror.make_env_model("PendulumKnown", None)
    i2c = mirror.I2cGraph(model, 100, g["Q"], g["R"], g["Qf"], 100, 0.0, g["mu_u"], g["sig_u"], None, None,
                          mirror.CubatureQuadrature(1, 0, 0))
    i2c.learn_msgs(60)
    policy_linear = mirror.TimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), 100, 1, 2)
    policy_linear.write(*i2c.get_local_linear_policy())
    np.random.seed(0)
    xs, ys, zs, zs_term = env.batch_eval(policy_linear, 10)
    assert len(xs) == 10 and xs[0].shape == (100, 3) and ys[0].shape == (100, 2) and zs[0].shape == (100, 4)
    assert zs_term[0].shape == (1, 3)
    # the closed loop follows the plan: realised trajectory close to the smoothed one
    plan = i2c.get_marginal_trajectory()
    assert np.max(np.abs(np.mean(xs, axis=0)[:, :2] - plan[:, :2])) < 0.5
    # host loop over the same controller reproduces a roll-out (same disturbances via the seed)
    np.random.seed(1)
    x1, y1, z1, zt1 = env.run(policy_linear)
    np.random.seed(1)
    eta = np.random.multivariate_normal(np.zeros(2), env.sig_eta, (1, 1, 100))[0, 0]
    x = np.asarray(model.x0, float)[:, 0].copy()
    from oracle import envs as E

    sys_ = E.Pendulum()
    for t in range(100):
        u = policy_linear(t, x[:, None])[:, 0]
        assert np.allclose(x1[t], np.concatenate((x, u)), rtol=0, atol=1e-9)
        x = sys_.dynamics(np.concatenate((x, u))) + eta[t]
    policy = mirror.ExpertTimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), 100, 1, 2, soft=False)
    policy.write(*i2c.get_local_expert_linear_policy())
    xs, ys, zs, zs_term = env.batch_eval(policy, 4, deterministic=False)
    assert np.all(np.isfinite(np.asarray(xs)))


def test_linearize_pendulum_flow(mirror):
    """scripts/i2c_run.py with experiments/pendulum_known.py (Linearize inference on a nonlinear env):
    the mirror runs the same calls; compared with the oracle (finite-difference Jacobians)."""
    from oracle import i2c_oracle as O

    rng = np.random.default_rng(2)
    T = 50
    mu_u = 1e-2 * rng.normal(size=(T, 1))
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    model = mirror.make_env_model("PendulumKnown", None)
    i2c = mirror.I2cGraph(model, T, Q, R, Q, 100.0, 0.5, mu_u, 2.0 * np.eye(1), None, None, mirror.Linearize())
    ref = O.make_graph("PendulumKnown", T, Q, R, Q, 100.0, 0.5, mu_u, 2.0 * np.eye(1), inference=O.Linearize())
    for _ in range(4):
        i2c.learn_msgs()
        ref.learn_msgs()
    assert relerr(np.array(i2c.alphas), np.array([a[0] for a in ref.alphas])) < 1e-9
    K, k, s = i2c.get_local_linear_policy()
    Kr, kr, sr = ref.get_local_linear_policy()
    # the oracle differentiates the dynamics by the complex-step method (exact to rounding, pinned to the unmodified reference by
    # tests/golden/*_linearize_*.npz), the kernel by forward-mode AD: round-off level agreement
    assert relerr(K, Kr[0]) < 1e-9 and relerr(k, kr[0]) < 1e-9 and relerr(s, sr[0]) < 1e-9


def test_alpha_helpers_match_device_update(mirror):
    """compute_update_alpha / calculate_alpha / get_z_covar (i2c.py:913-992) recomputed on the host from the device
    messages reproduce the alpha the M-step kernel produced; plot_* calls of the scripts are accepted."""
    g = golden("pendulum_T200_x0pert")
    T = 60
    model = mirror.make_env_model("PendulumKnown", None)
    model.x0 = g["x0"].reshape(-1, 1)
    graph = mirror.I2cGraph(model, T, g["Q"], g["R"], g["Qf"], 100.0, 0.5, g["mu_u"][:T], g["sig_u"], None, None,
                            mirror.CubatureQuadrature(1, 0, 0))
    graph._forward_backward_msgs()
    a0 = graph.alpha
    desired = graph.calculate_alpha(graph.get_z_covar(), graph.get_z_terminal_covar())
    graph.compute_update_alpha(True)  # host recomputation + i2c_set_alpha
    host_alpha = graph.alpha
    assert abs(graph.alphas_desired[-1] - desired) < 1e-12 * desired
    # same sweep through the fused device path
    graph2 = mirror.I2cGraph(model, T, g["Q"], g["R"], g["Qf"], 100.0, 0.5, g["mu_u"][:T], g["sig_u"], None, None,
                             mirror.CubatureQuadrature(1, 0, 0))
    graph2.learn_msgs()
    assert abs(graph2.alphas_desired[-1] - desired) < 1e-10 * desired
    assert abs(graph2.alpha - host_alpha) < 1e-10 * host_alpha and host_alpha != a0
    assert graph.propagate_cost_improved is True
    graph.plot_metrics(0, 0, None, "msg")  # scripts/i2c_run.py:123
    graph.plot_traj(0, dir_name=None, filename="lqr")  # scripts/lqr_compare.py:172
    with pytest.raises(AttributeError):
        graph.no_such_attribute


def test_likelihood_diagnostics(mirror):
    """I2cGraph.calc_likelihood (i2c.py:1135-1164) after every EM iteration, against the reference (likelihood_kat.npz);
    the values are sums of ~1e5-sized terms, hence the relative tolerance."""
    g = golden("likelihood_kat")
    model = mirror.make_env_model("PendulumKnown", None)
    graph = mirror.I2cGraph(model, int(g["T"]), g["Q"], g["R"], g["Qf"], float(g["alpha0"]), float(g["tol"]), g["mu_u"],
                            g["sig_u"], None, None, mirror.CubatureQuadrature(1, 0, 0))
    for _ in range(4):
        graph.learn_msgs()
        graph.calc_likelihood()
    for name in ("likelihoods", "likelihoods_xu", "likelihoods_z", "risk"):
        assert relerr(np.asarray(getattr(graph, name), float).reshape(-1), g[name]) < 1e-7, name
    vals = [3.0, 2.5, 2.0, 2.2, 1.9, 1.8, 1.7]
    mine = np.array([[float(x) if x is not None else -1.0 for x in
                      (graph.list_minima(vals[:n], 2, 2), graph.list_minima(vals[:n], 2, 3))] for n in range(1, 8)])
    assert np.array_equal(mine, g["minima"])
