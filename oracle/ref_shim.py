"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED upstream reference.

This module makes the reference tree (``/root/reference`` in the build container, or its
byte-for-byte copy ``oracle/_ref`` made by ``oracle/install_ref.py``, which is what travels to
the GPU box) importable on this image's NumPy 2.x without touching its sources (recipe:
SURVEY.md Appendix C).  Users: ``tests/golden/make_golden.py`` (golden vectors), the
reference arm of ``bench.py`` (``--impl reference`` / ``cpu_baseline``: the reference's own
``I2cGraph.learn_msgs`` timed on the host cores) and ``tests/test_dropin_scripts.py``.
Nothing in the product path may import this file.

What it does, before any ``import i2c``:
  1. restores the NumPy aliases the reference still uses
     (np.asscalar, np.NINF, np.Inf, np.float);
  2. installs permissive stub modules for the plotting / gym / autograd
     dependencies that are absent here and are never called on the numeric path;
  3. puts /root/reference (and /root/reference/scripts) on sys.path.
"""
import importlib
import os
import sys
import types

import numpy as np

def _default_root():
    """$I2C_REFERENCE_ROOT, else the byte-for-byte copy oracle/install_ref.py put under oracle/_ref/ (the only one that
    exists on the GPU box), else the build container's /root/reference."""
    env = os.environ.get("I2C_REFERENCE_ROOT")
    if env:
        return env
    local = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isdir(os.path.join(local, "i2c")) or not os.path.isdir("/root/reference/i2c"):
        return local
    return "/root/reference"


REFERENCE_ROOT = _default_root()


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub; calling it is a no-op."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return _Stub(self.__name__ + "()")

    def __iter__(self):
        return iter(())


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "i2c"))


def install(paths=True, extra_stubs=()):
    """Idempotently install aliases, stubs and (paths=True) the sys.path entries that make ``import i2c`` resolve to the
    REFERENCE.  paths=False only installs the NumPy aliases and the stubs: used by tests/test_dropin_scripts.py to run the
    reference's scripts against this repo's ``i2c`` package instead."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in extra_stubs:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: np.asarray(a).item()
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if not hasattr(np, "float"):
        np.float = float

    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.patches",
        "tikzplotlib",
        "matplotlib2tikz",
        "gym",
        "gym.spaces",
        "imageio",
        "pygifsicle",
    ):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)

    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")
        ag.numpy = np

        def jacobian(fn, argnum=0):
            """Stand-in for ``autograd.jacobian`` (env_autograd.py:22,57,170), which is absent from this image: the
            complex-step derivative Im f(x + i h e_j) / h with h = 1e-30.  For the reference's dynamics functions (sums,
            products, sin / cos, np.power, np.linalg.inv, np.clip away from the limits -- all analytic) it has no
            truncation and no cancellation error: it equals the exact derivative, which is what autograd returns, to
            rounding.  The reference's OWN dynamics functions are differentiated, unmodified.  Result shape as
            autograd's: ``fn(x).shape + x.shape``."""
            def jac(*args):
                x = np.asarray(args[argnum], dtype=float)
                y0 = np.asarray(fn(*args))
                out = np.empty(y0.shape + x.shape)
                h = 1e-30
                for idx in np.ndindex(*x.shape):
                    xc = x.astype(complex)
                    xc[idx] += 1j * h
                    a = list(args)
                    a[argnum] = xc
                    out[(Ellipsis,) + idx] = np.imag(np.asarray(fn(*a))) / h
                return out

            return jac

        ag.jacobian = jacobian
        sys.modules["autograd"] = ag
        sys.modules["autograd.numpy"] = np

    if "numdifftools" not in sys.modules:
        nd = types.ModuleType("numdifftools")

        class Jacobian:  # lazily evaluated, only touched by dead Furuta code
            def __init__(self, fn, *a, **k):
                self.fn = fn

            def __call__(self, *a, **k):
                raise NotImplementedError("numdifftools is not available")

        nd.Jacobian = Jacobian
        sys.modules["numdifftools"] = nd

    if paths:
        for p in (os.path.join(REFERENCE_ROOT, "scripts"), REFERENCE_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)


def _purge(prefixes):
    for name in list(sys.modules):
        if any(name == p or name.startswith(p + ".") for p in prefixes):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", None) or ""
            # only purge modules that are not from the reference tree
            if not f.startswith(REFERENCE_ROOT):
                del sys.modules[name]


def load():
    """Return the reference's modules as a namespace (i2c, quadrature, ...)."""
    install()
    # a different package called ``i2c`` (this repo's drop-in mirror) may already be
    # imported in the same interpreter: make sure we get the reference's.
    _purge(["i2c", "experiments"])
    ns = types.SimpleNamespace()
    ns.i2c = importlib.import_module("i2c.i2c")
    ns.quadrature = importlib.import_module("i2c.inference.quadrature")
    ns.exp_types = importlib.import_module("i2c.exp_types")
    ns.model = importlib.import_module("i2c.model")
    ns.env_def = importlib.import_module("i2c.env_def")
    ns.utils = importlib.import_module("i2c.utils")
    ns.mpc = importlib.import_module("i2c.policy.mpc")
    assert ns.i2c.__file__.startswith(REFERENCE_ROOT), ns.i2c.__file__
    return ns


def unload():
    """Drop the reference's modules so this repo's ``i2c`` mirror can be imported."""
    for name in list(sys.modules):
        if name == "i2c" or name.startswith("i2c.") or name.startswith("experiments"):
            f = getattr(sys.modules[name], "__file__", None) or ""
            if f.startswith(REFERENCE_ROOT):
                del sys.modules[name]
    for p in (os.path.join(REFERENCE_ROOT, "scripts"), REFERENCE_ROOT):
        while p in sys.path:
            sys.path.remove(p)


def load_experiment(name, seed):
    """Import scripts/experiments/<name>.py with the global seed set just before
    (the reference draws mu_u at import time: i2c_run.py:215-217)."""
    install()
    np.random.seed(seed)
    modname = f"experiments.{name}"
    if modname in sys.modules:
        del sys.modules[modname]
    return importlib.import_module(modname)
