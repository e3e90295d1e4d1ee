import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "input-inference-for-control_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def relerr(a, ref, floor=0.0):
    """Norm-wise relative error max|a-ref| / max(max|ref|, floor) over the whole tensor
    (element-wise relative error is meaningless, e.g. K[T-1] ~ 1e-16;
    ``floor`` guards tensors that are pure round-off noise in the reference itself)."""
    a, ref = np.asarray(a, float), np.asarray(ref, float)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if not np.all(np.isfinite(a)):
        return float("inf")
    denom = max(float(np.max(np.abs(ref))) if ref.size else 0.0, floor)
    if denom == 0.0:
        e = float(np.max(np.abs(a))) if a.size else 0.0
    else:
        e = float(np.max(np.abs(a - ref)) / denom)
    _report(e, floor)
    return e


def _report(e, floor):
    """With I2C_PARITY_REPORT=<file> every measured error is appended as {test, where, err}: the evidence behind the
    tolerances written in the tests (tools/parity_floors.py summarises it; profiles/*parity_floors*)."""
    path = os.environ.get("I2C_PARITY_REPORT")
    if not path:
        return
    import inspect
    import json

    fr = inspect.stack()[2]
    where = f"{os.path.basename(fr.filename)}:{fr.lineno}"
    if os.path.basename(fr.filename) == "conftest.py" or fr.function in ("rel", "compare_cells"):
        up = inspect.stack()[3]
        where += f"<{os.path.basename(up.filename)}:{up.lineno}"
    with open(path, "a") as f:
        f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "where": where + ("[gain]" if floor > 0 else ""), "err": e}) + "\n")


# tensors whose value is a solve against a small covariance: their round-off floor in the reference
# itself is 1e-10 .. 1e-8 (cancellation in sum_p w x y^T - m m^T divided by Sigma ~ 1e-5), see DESIGN.md
GAINS = ("J_dyn", "K", "k")


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
