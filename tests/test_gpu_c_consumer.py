"""The C-ABI used from plain C (tests/c_abi/consumer.c, compiled with gcc against include/i2c_b200.h and linked to
lib/libi2c_b200.so -- no Python, no torch in that process): same controllers and alpha schedule as the Python host path on the
same inputs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def lcg_stream(n, seed=12345):
    out = np.empty(n)
    s = seed
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out[i] = ((s >> 11) & ((1 << 53) - 1)) / float(1 << 53) - 0.5
    return out


def build_consumer(tmp_path):
    libdir = os.path.join(ROOT, "input-inference-for-control_b200", "lib")
    exe = os.path.join(str(tmp_path), "consumer")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "consumer.c"), "-L" + libdir, "-li2c_b200",
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    return exe


def test_c_consumer_matches_python_host(i2c_b200, tmp_path):
    import shutil

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box")
    B, T, n_iter = 200, 40, 4
    exe = build_consumer(tmp_path)
    r = subprocess.run([exe, str(B), str(T), str(n_iter)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == f"ok B={B} T={T} n_iter={n_iter} failed=0"
    got = {ln.split()[0]: float(ln.split()[1]) for ln in lines[1:]}
    # the same problem through the Python host classes
    u = lcg_stream(2 * B + B * T)
    x0 = np.stack((np.pi + 0.6 * u[0:2 * B:2], 1.0 * u[1:2 * B:2]), axis=1)
    mu_u = (0.02 * u[2 * B:]).reshape(B, T, 1)
    Q, R = np.diag([1.0, 100.0, 1.0]), np.diag([2.0])
    g = i2c_b200.BatchedI2c("PendulumKnown", B, T, Q, R, Q, 100.0, 0.0, mu_u, 2.0 * np.eye(1), x0=x0,
                            sig_x0=1e-5 * np.eye(2), sig_eta=1e-5 * np.eye(2), max_iters=16)
    g.learn(n_iter)
    K, k, s = g.get_local_linear_policy()
    w7 = 1.0 + (np.arange(K.size) % 7)
    w5 = 1.0 + (np.arange(k.size) % 5)
    ref = {"K": float(np.sum(K.ravel() * w7)), "k": float(np.sum(k.ravel() * w5)), "sigK": float(np.sum(s)),
           "alpha": float(np.sum(np.array(g.metrics["alpha"])))}
    for key in ref:
        assert abs(got[key] - ref[key]) <= 1e-11 * max(1.0, abs(ref[key])), (key, got[key], ref[key])
