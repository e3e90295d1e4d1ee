#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- puts a copy of the UNMODIFIED upstream reference under oracle/_ref/.

    python oracle/install_ref.py            # from /root/reference (or $I2C_REFERENCE_SRC)

The reference is a pure-Python package (setup.py: install_requires=["numpy"]); its `i2c/` package and the `scripts/` tree
(experiment definitions + the runnable scripts of BASELINE configs 1, 2, 4, 5) are copied byte for byte -- nothing is
built or patched; oracle/ref_shim.py makes the tree importable on this image's NumPy 2.x from the outside.
oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so
it travels to the GPU box, where it serves as
  * the CPU arm of the benchmark: bench.py --impl reference times the reference's own I2cGraph.learn_msgs
    (cpu_baseline.kind = "reference"),
  * the drop-in check: tests/test_dropin_scripts.py runs the reference's scripts against the CUDA mirror package.
Nothing in the product path (input-inference-for-control_b200/) reads this directory.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("I2C_REFERENCE_SRC", "/root/reference")
KEEP = ("i2c", "scripts", "baselines", "setup.py", "README.md", "requirements.txt", "LICENSE")


def tree_digest(root):
    h = hashlib.sha256()
    for dp, dn, fn in sorted(os.walk(root)):
        dn.sort()
        for f in sorted(fn):
            if f.endswith(".pyc"):
                continue
            p = os.path.join(dp, f)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def install(force=False):
    """Idempotent; returns the destination, or None when no reference source tree is available (GPU box)."""
    if not os.path.isdir(os.path.join(SRC, "i2c")):
        return DST if os.path.isdir(os.path.join(DST, "i2c")) else None
    stamp = os.path.join(DST, ".source_sha256")
    if not force and os.path.exists(stamp):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in KEEP:
        s = os.path.join(SRC, name)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DST, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "_results"))
        elif os.path.exists(s):
            shutil.copy2(s, os.path.join(DST, name))
    with open(stamp, "w") as f:
        f.write(tree_digest(os.path.join(DST, "i2c")) + "\n")
    return DST


if __name__ == "__main__":
    d = install(force="--force" in sys.argv)
    print(d if d else "no reference source tree at " + SRC)
