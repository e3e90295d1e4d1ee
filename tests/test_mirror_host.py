"""Host-side pieces of the drop-in mirror that need no GPU: the trajectory-cost evaluators that consume the roll-outs in
scripts/i2c_run.py:66-106 and the .npy result files they write.  Golden = the reference's
own classes run in the build container (tests/golden/make_golden.py eval)."""
import os

import numpy as np

from conftest import golden, relerr


def test_evaluators_match_reference(tmp_path):
    from i2c.utils import StochasticTrajectoryEvaluator, TrajectoryEvaluator

    g = golden("evaluator_kat")
    ev = StochasticTrajectoryEvaluator(g["W"], g["Wf"], g["sg"], g["sg_term"], int(g["dim_x"]))
    det = TrajectoryEvaluator(g["W"], g["Wf"], g["sg"], g["sg_term"], int(g["dim_x"]))
    for k in range(3):
        ev.eval(g[f"{k}/trajs"], g[f"{k}/terms"], g[f"{k}/plan"], g[f"{k}/plan_term"])
        det.eval(g[f"{k}/trajs"][0], g[f"{k}/terms"][:1], g[f"{k}/plan"], g[f"{k}/plan_term"])
    for name in ["mu_actual_cost", "max_actual_cost", "min_actual_cost", "actual_cost_10", "actual_cost_90", "planned_cost"]:
        assert relerr(np.asarray(getattr(ev, name)), g[f"stoch/{name}"]) < 1e-13, name
    assert relerr(np.asarray(det.actual_cost), g["det/actual_cost"]) < 1e-13
    assert relerr(np.asarray(det.planned_cost), g["det/planned_cost"]) < 1e-13
    ev.save("x", str(tmp_path))
    det.save("y", str(tmp_path))
    written = sorted(os.listdir(tmp_path))
    assert written == sorted(k[5:] for k in g.files if k.startswith("file/"))  # same file names as the reference
    for f in written:
        a, b = np.load(tmp_path / f), g[f"file/{f}"]
        assert a.dtype == b.dtype and a.shape == b.shape and relerr(a, b) < 1e-13, f
