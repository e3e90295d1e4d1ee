"""Two handles on two different GPUs in ONE process (every handle entry point must run on the handle's own
device whatever device is current, and the > 48 KB shared-memory opt-in is per device).  Needs >= 2 visible GPUs."""
import numpy as np
import pytest

from conftest import relerr
from test_gpu_parity import i2c_b200  # noqa: F401
from tools_inputs import make_case

pytestmark = pytest.mark.gpu


def test_two_handles_on_two_devices(i2c_b200):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    graphs = [make_case(i2c_b200, "CartpoleKnown", 96, 20, device=d) for d in (0, 1)]  # cart-pole: > 48 KB of dynamic smem
    torch.cuda.set_device(0)  # the current device stays 0 throughout: handle 1 must guard itself
    for g in graphs:
        g.learn(3)
    for g in graphs:
        assert np.all(g.status()[0] == 0)
        K, k, s = g.get_local_linear_policy()
        outs.append((K, k, s, g.field("mu_xu0_m"), g.alpha))
    assert torch.cuda.current_device() == 0
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)  # same inputs, same code: bit-identical on both devices
    snap = graphs[1].snapshot()
    graphs[1].learn(1)
    graphs[1].restore(snap)
    assert relerr(graphs[1].field("mu_xu0_m"), outs[1][3]) == 0.0
    for g in graphs:
        g.close()
