#!/usr/bin/env python
"""A/B builds for kernel experiments: recompile the translation units of the given environments with extra -D flags and
link them with the current objects of every other unit into lib/ab/<name>.so (git-ignored, travels to the GPU box).
Select it at run time with I2C_B200_LIB=<path>.      python tools/ab_build.py <name> [--envs 2,4] [-DFLAG ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    name = sys.argv[1]
    envs, defs = [2], []
    it = iter(sys.argv[2:])
    for a in it:
        if a == "--envs":
            envs = [int(x) for x in next(it).split(",")]
        else:
            defs.append(a)
    ge.build()
    out = os.path.join(ge.LIBDIR, "ab", name)
    os.makedirs(out, exist_ok=True)
    objs, procs = [], []
    for src, obj, d in ge.UNITS:
        is_env = src == "i2c_env_inst.cu"
        k = int(d[0].split("=")[1]) if is_env else -1
        if is_env and k in envs:
            o = os.path.join(out, obj)
            cmd = [ge.NVCC, *ge.ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "550,177", *d,
                   *defs, "-c", os.path.join(ge.CSRC, src), "-o", o]
            procs.append(subprocess.Popen(cmd))
            objs.append(o)
        else:
            objs.append(os.path.join(ge.LIBDIR, obj))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    lib = os.path.join(ge.LIBDIR, "ab", name + ".so")
    subprocess.check_call([ge.NVCC, *ge.ARCH, "-shared", "-o", lib, *objs])
    print(lib)


if __name__ == "__main__":
    main()
