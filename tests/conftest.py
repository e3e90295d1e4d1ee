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
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - ref)) / denom)


# tensors whose value is a solve against a small covariance: their round-off floor in the reference
# itself is 1e-10 .. 1e-8 (cancellation in sum_p w x y^T - m m^T divided by Sigma ~ 1e-5), see DESIGN.md
GAINS = ("J_dyn", "K", "k")


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
