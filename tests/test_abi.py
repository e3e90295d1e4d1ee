"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/i2c_b200.h declares; compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    import i2c_b200

    return i2c_b200


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "i2c_b200.h")).read()
    names = set(re.findall(r"\b(i2c_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    lib = ctypes.CDLL(built.capi.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(built.capi.EXPORTS) == names


def test_env_dims_and_workspace(built):
    capi = built.capi
    assert capi.env_dims(capi.ENV_IDS["PendulumKnown"]) == (2, 1, 4, 3, 0, 0)
    assert capi.env_dims(capi.ENV_IDS["DoubleCartpoleKnown"]) == (6, 1, 9, 8, 0, 0)
    assert capi.env_dims(capi.ENV_IDS["Quadrotor"]) == (6, 2, 8, 6, 0, 8)
    assert capi.env_dims(capi.ENV_IDS["LinearKnown"]) == (2, 1, 3, 2, 8, 0)
    cfg = capi.Config(capi.ABI_VERSION, capi.ENV_IDS["PendulumKnown"], 0, 4096, 200, 128, 0, 0, 0, 1.0, 0.0, 0.0)
    n = ctypes.c_size_t()
    capi.check(capi.lib().i2c_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)))
    # prior + post (13 el.) + filtered (20 el.) records of 4096 x 200 cells plus staging
    assert n.value > 4096 * 200 * (13 + 13 + 20) * 8
    cfg.abi_version = 99
    assert capi.lib().i2c_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)) != 0
    assert b"ABI" in capi.lib().i2c_last_error()


def test_host_env_constants_match_device_dims(built):
    capi, envs = built.capi, built.envs
    for name, eid in capi.ENV_IDS.items():
        e = envs.make(name)
        dx, du, dz, dzt, npar, dy = capi.env_dims(eid)
        assert (e.dim_x, e.dim_u, e.dim_z, e.dim_z_term, e.dim_y) == (dx, du, dz, dzt, dy)
        assert e.x0.shape == (dx,) and e.sig_x0.shape == (dx, dx) and e.sig_eta.shape == (dx, dx)
        assert e.zg.shape == (dz,) and e.zg_term.shape == (dzt,)


def test_no_cpu_fallback(built, has_cuda):
    if has_cuda:
        pytest.skip("GPU present")
    with pytest.raises(built.I2cError):
        built.BatchedI2c("PendulumKnown", 4, 10, np.eye(3), np.eye(1), np.eye(3), 1.0, 0.0, np.zeros((10, 1)), np.eye(1))
    with pytest.raises(built.I2cError):
        built.quadrature("PendulumKnown", "observe", np.zeros((1, 3)), np.eye(3))
    with pytest.raises(KeyError):
        built.envs.make("FurutaKnown")


def test_header_is_c99_and_c_consumer_links(built, has_cuda, tmp_path):
    """include/i2c_b200.h is plain C (gcc -std=c99 -pedantic), and a C program (tests/c_abi/consumer.c) links against the
    library without Python / torch; without a GPU its i2c_create is refused loudly (exit code 3): no CPU fallback."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box")
    libdir = os.path.dirname(built.capi.LIB_PATH)
    exe = os.path.join(str(tmp_path), "consumer")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "consumer.c"), "-L" + libdir, "-li2c_b200",
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    r = subprocess.run([exe, "8", "10", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if has_cuda:
        assert r.returncode == 0 and r.stdout.startswith("ok B=8 T=10 n_iter=2 failed=0"), r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stdout, r.stdout + r.stderr
