"""Drop-in check (BASELINE north star: "scripts/i2c_run.py, the LQR comparison and the MPC scripts drop onto the
new path unchanged"): the reference's scripts are executed UNMODIFIED -- their sources are read from the reference tree
(oracle/_ref, the byte-for-byte copy made by oracle/install_ref.py; /root/reference in the build container) -- once with the
reference's own ``i2c`` package (CPU) and once with this repo's CUDA mirror package, and what they computed is compared.
Both runs happen in fresh interpreters (tests/dropin_runner.py).  The absent plotting / gym / Box2D modules are stubbed from
the outside on both sides; the Box2D quadrotor step is the fp64 restatement on both sides (parity unpinned)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr

RUNNER = os.path.join(ROOT, "tests", "dropin_runner.py")
SCRIPTS = ["i2c_run", "lqr_compare", "nonlinear_covariance_control", "mpc_quad"]


def have_reference():
    sys.path.insert(0, ROOT)
    from oracle import ref_shim

    return ref_shim.available()


needs_ref = pytest.mark.skipif(not have_reference(), reason="reference tree (oracle/_ref or /root/reference) not present")


def run(impl, script, out, *extra):
    cmd = [sys.executable, RUNNER, "--impl", impl, "--script", script, "--out", out, *extra]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


@needs_ref
@pytest.mark.parametrize("script", SCRIPTS)
def test_scripts_import_against_the_mirror(script, tmp_path):
    """No GPU needed: every name the scripts import from ``i2c`` exists in the mirror package (i2c.env_def.BaseDef,
    i2c.model.BaseModelKnown, i2c.i2c.PLOT_TIKZ, i2c.utils.make_results_folder / configure_plots / covariance_2d / ...)."""
    out = run("mirror", script, str(tmp_path / "x.npz"), "--import-only")
    assert "import ok" in out


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("script,extra,tol_s,tol_g", [
    ("i2c_run", ["--iters", "30"], 1e-9, 2e-9),            # scripts/i2c_run.py:run() on experiments/pendulum_known_quad.py
    ("lqr_compare", [], 5e-9, 1e-4),                        # gains after the Riccati messages: see tests/test_gpu_lqr.py
    ("nonlinear_covariance_control", [], 5e-8, 5e-8),       # pendulum act-reg covariance control (floor 2e-9 in the reference)
])
def test_script_outputs_match_the_reference(script, extra, tol_s, tol_g, tmp_path):
    ref, mine = str(tmp_path / "ref.npz"), str(tmp_path / "mine.npz")
    run("reference", script, ref, *extra)
    run("mirror", script, mine, *extra)
    a, b = np.load(ref), np.load(mine)
    assert int(a["n_graphs"]) == int(b["n_graphs"]) >= 1
    for key in a.files:
        if key in ("n_graphs", "files"):
            continue
        gain = key.endswith("/K") or key.endswith("/k")
        assert relerr(b[key], a[key], 1e-6 if gain else 0.0) < (tol_g if gain else tol_s), key
    if script == "i2c_run":
        # result files of the run: trajectories of save_traj / save_trajectories, evaluator costs
        need = {"xu_plan.npy", "x_plan.npy", "u_plan.npy", "z_plan.npy", "xu_real.npy", "dx_real.npy", "x_real.npy", "u_real.npy",
                "cost_actual_mean_iter.npy", "cost_plan_iter.npy", "cost_actual_mean_episodic.npy", "cost_plan_episodic.npy"}
        assert need <= set(b["files"].tolist()), set(b["files"].tolist())
        assert need <= set(a["files"].tolist())


@needs_ref
@pytest.mark.gpu
def test_mpc_quad_single_experiment_matches_the_reference(tmp_path):
    """scripts/mpc_state_est/mpc_quad.py: single_experiment(use_i2c=True, feedforward=False, low_noise=True) -- 100 closed-loop
    control steps with the script's own QuadrotorDef / QuadrotorKnown(QuadrotorDef, BaseModelKnown) classes, gym-style
    simulator and global-RNG noise; result files state_*.npy / obs_*.npy / <name>.npy as process_results.py reads them."""
    ref, mine = str(tmp_path / "ref.npz"), str(tmp_path / "mine.npz")
    run("reference", "mpc_quad", ref)
    run("mirror", "mpc_quad", mine)
    a, b = np.load(ref), np.load(mine)
    # closed loop over 100 steps: the first steps agree to round-off, the end to the loop's amplification
    assert relerr(b["states"][:10], a["states"][:10]) < 1e-9
    assert relerr(b["obs"][:10], a["obs"][:10]) < 1e-9
    assert relerr(b["states"], a["states"]) < 1e-6
    assert abs(float(b["cost"]) - float(a["cost"])) < 1e-6 * abs(float(a["cost"]))
