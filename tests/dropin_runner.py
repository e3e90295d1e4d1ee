"""Runs ONE of the reference's scripts, unmodified, in this process -- against the reference's own ``i2c`` package
(--impl reference, CPU) or against this repo's CUDA mirror package (--impl mirror) -- and dumps what the script computed.
Spawned twice per script by tests/test_dropin_scripts.py, which compares the two dumps.  Test infrastructure only.

The script sources come from the reference tree (oracle/_ref, or /root/reference): nothing is patched in the files; what
the image lacks is stubbed from the outside exactly as oracle/ref_shim.py does (matplotlib, gym, Box2D, tikzplotlib, ...),
and the Box2D rigid-body step of the quadrotor definition is replaced by the fp64 restatement of oracle/envs.py on BOTH
sides (Box2D is absent: SURVEY.md 8c).  The inference graphs the script builds are captured by wrapping I2cGraph.__init__.
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "input-inference-for-control_b200")
sys.path.insert(0, ROOT)


class Dummy(object):
    """Stands in for any object of the absent plotting / rendering / physics libraries: every call, attribute, item access
    and iteration is accepted and does nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Dummy()

    def __getitem__(self, i):
        return Dummy()

    def __setitem__(self, i, v):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0


class StubModule(__import__("types").ModuleType):
    """Module whose CapitalisedNames (and everything of Box2D) are the Dummy class -- usable as a base class, e.g.
    ``class ContactDetector(contactListener)``, ``class Quadrotor(gym.Env)`` -- and whose other names are Dummy instances."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name == "subplots":
            def subplots(nrows=1, ncols=1, *a, **k):
                shape = tuple(n for n in (nrows, ncols) if n > 1)
                ax = np.empty(shape, dtype=object)
                for i in np.ndindex(*shape):
                    ax[i] = Dummy()
                return Dummy(), (ax if shape else Dummy())
            return subplots
        if name[:1].isupper() or self.__name__.startswith("Box2D"):
            return Dummy
        return Dummy()


def setup(impl):
    from oracle import ref_shim

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "tikzplotlib", "matplotlib2tikz", "Box2D", "Box2D.b2",
                 "gym", "gym.spaces", "imageio", "pygifsicle", "baselines", "baselines.ilqr", "trajopt", "git"):
        sys.modules[name] = StubModule(name)
    for parent, child in (("matplotlib", "pyplot"), ("matplotlib", "patches"), ("Box2D", "b2"), ("gym", "spaces"),
                          ("baselines", "ilqr")):
        setattr(sys.modules[parent], child, sys.modules[f"{parent}.{child}"])
    ref_shim.install(paths=False)
    ref = ref_shim.REFERENCE_ROOT
    # `scripts.experiments.x` / `experiments.x` / `baselines` come from the reference tree on both sides; `i2c` from the
    # reference (impl reference) or from this repo (impl mirror)
    paths = [os.path.join(ref, "scripts"), ref]
    if impl == "mirror":
        sys.path[:0] = [PKG] + paths
        import i2c

        assert os.path.abspath(i2c.__file__).startswith(PKG), i2c.__file__
    else:
        sys.path[:0] = [ref, os.path.join(ref, "scripts")]
        import i2c

        assert os.path.abspath(i2c.__file__).startswith(os.path.abspath(ref)), i2c.__file__
    return ref


DUMP = {}


def capture_graphs():
    import i2c.i2c as mod

    made = []
    orig = mod.I2cGraph.__init__
    orig_close = mod.I2cGraph.close

    def init(self, *a, **k):
        orig(self, *a, **k)
        made.append(self)

    def close(self):  # scripts/i2c_run.py:160 closes the graph: take the dump first (the mirror frees the device state)
        graph_dump(self, DUMP, f"g{made.index(self)}")
        return orig_close(self)

    mod.I2cGraph.__init__ = init
    mod.I2cGraph.close = close
    # figures are outside the path, and the reference's own plot_traj raises on the Linearize path (mu_xu0_f_prev is never
    # set there: SURVEY.md section 2 "known-broken reference code"): no-ops on both sides (the mirror's already are)
    for name in dir(mod.I2cGraph):
        if name.startswith("plot_"):
            setattr(mod.I2cGraph, name, lambda self, *a, **k: None)
    return made


def graph_dump(g, d, tag):
    try:
        K, k, s = g.get_local_linear_policy()
    except AttributeError:  # MPC graphs: freshly appended cells have no controller yet (policy/mpc.py:174-176)
        return
    d[f"{tag}/K"], d[f"{tag}/k"], d[f"{tag}/sigK"] = np.asarray(K, float), np.asarray(k, float), np.asarray(s, float)
    d[f"{tag}/alphas"] = np.asarray(g.alphas, float).reshape(-1)
    d[f"{tag}/costs_m"] = np.asarray(g.costs_m, float).reshape(-1)
    d[f"{tag}/xu"] = np.asarray(g.get_marginal_state_action(), float).reshape(len(g.cells), -1)
    if getattr(g, "costs_pf", None):
        d[f"{tag}/costs_pf"] = np.asarray(g.costs_pf, float).reshape(-1)
    if getattr(g, "kl_terms", None):
        d[f"{tag}/kl_terms"] = np.asarray(g.kl_terms, float).reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["reference", "mirror"])
    ap.add_argument("--script", required=True, choices=["i2c_run", "lqr_compare", "nonlinear_covariance_control", "mpc_quad"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--iters", type=int, default=0, help="override N_INFERENCE of the experiment module (0 = as shipped)")
    ap.add_argument("--import-only", action="store_true", help="only import the script module (no CUDA device needed)")
    a = ap.parse_args()
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    setup(a.impl)
    import importlib

    if a.import_only:
        from oracle import ref_shim

        sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "mpc_state_est"))
        mod = importlib.import_module(a.script)
        for name in {"i2c_run": ["run", "save_trajectories"], "lqr_compare": ["main"], "nonlinear_covariance_control": ["main"],
                     "mpc_quad": ["single_experiment", "QuadrotorKnown", "QuadrotorDef"]}[a.script]:
            assert hasattr(mod, name), name
        print("dropin_runner: import ok", a.impl, a.script)
        return
    made = capture_graphs()
    d = {}
    work = tempfile.mkdtemp(prefix="dropin_")
    os.chdir(work)  # the scripts write _results/... relative to the working directory
    if a.script == "i2c_run":
        np.random.seed(0)  # i2c_run.py:215 set_seed(args.random_seed) before the experiment module is imported
        exp = importlib.import_module("experiments.pendulum_known_quad")
        if a.iters:
            exp.N_INFERENCE = a.iters
        mod = importlib.import_module("i2c_run")
        mod.N_EVAL = 2  # fewer evaluation roll-outs per iteration (they do not feed back into the inference)
        res_dir = os.path.join(work, "res")
        os.makedirs(res_dir)
        mod.run(exp, res_dir, None)
        d["files"] = np.array(sorted(os.listdir(res_dir)))
        for f in ("xu_plan.npy", "x_plan.npy", "u_plan.npy", "z_plan.npy"):
            d[f"file/{f}"] = np.load(os.path.join(res_dir, f))
    elif a.script == "lqr_compare":
        mod = importlib.import_module("lqr_compare")
        mod.main()
    elif a.script == "nonlinear_covariance_control":
        mod = importlib.import_module("nonlinear_covariance_control")
        mod.main()
    else:
        from oracle import envs as oenvs
        from oracle import ref_shim

        sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "mpc_state_est"))
        mq = importlib.import_module("mpc_quad")
        q = oenvs.Quadrotor()
        # Box2D is absent: the rigid-body step is the fp64 restatement on both sides (DESIGN.md "parity unpinned")
        mq.QuadrotorDef.init_world = lambda self: None
        mq.QuadrotorDef.gravity = property(lambda self: q.gravity)
        mq.QuadrotorDef.step = lambda self, x, u: oenvs.Quadrotor.dynamics(
            np.concatenate((np.asarray(x, float).reshape(-1), np.clip(np.asarray(u, float).reshape(-1), 0.0, 30.0)))[None, :])[0]
        # names the script binds under `if __name__ == "__main__":` (mpc_quad.py:725-735)
        import scipy.linalg as la
        from os.path import dirname, exists, join, realpath

        import matplotlib.pyplot as plt
        from i2c.i2c import I2cGraph
        from i2c.policy.mpc import PartiallyObservedMpcPolicy

        for name, val in dict(la=la, join=join, exists=exists, realpath=lambda p: os.path.join(work, "x.py"), dirname=dirname,
                              plt=plt, I2cGraph=I2cGraph, PartiallyObservedMpcPolicy=PartiallyObservedMpcPolicy).items():
            setattr(mq, name, val)
        mq.single_experiment(use_i2c=True, feedforward=False, low_noise=True, seed=3, name="i2c_FB_low_3")
        res = os.path.join(work, "_results")
        d["states"] = np.load(os.path.join(res, "state_i2c_FB_low_3.npy"))
        d["obs"] = np.load(os.path.join(res, "obs_i2c_FB_low_3.npy"))
        d["cost"] = np.load(os.path.join(res, "i2c_FB_low_3.npy"))
    d["n_graphs"] = len(made)
    for i, g in enumerate(made):
        if f"g{i}/alphas" not in DUMP:
            graph_dump(g, DUMP, f"g{i}")
    d.update(DUMP)
    np.savez(a.out, **d)
    print("dropin_runner: ok", a.impl, a.script, "graphs:", len(made))


if __name__ == "__main__":
    main()
