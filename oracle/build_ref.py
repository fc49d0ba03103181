#!/usr/bin/env python
"""Build the UNMODIFIED reference deposition module into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may use it, and only as the checker or the CPU baseline.

What it does: runs Cython on the reference's own source *where it lies*
(/root/reference/postpic/particles/_particlestogrid.pyx -- the one native file
of postpic, setup.py:31) and compiles the generated C with the same flags the
reference's setup.py ends up with (plain setuptools defaults: -O2, numpy
include dir, no -march, no -ffast-math, no OpenMP).  Outputs go ONLY to
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  No reference
source is copied into the repository; the intermediate C file is deleted.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODNAME = "_particlestogrid"


def so_path():
    return os.path.join(OUT, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(reference="/root/reference", force=False, verbose=True):
    """Returns the path of the built extension, or None when the reference tree is absent."""
    pyx = os.path.join(reference, "postpic", "particles", MODNAME + ".pyx")
    target = so_path()
    if os.path.exists(target) and not force:
        return target
    if not os.path.exists(pyx):
        return None
    import numpy
    os.makedirs(OUT, exist_ok=True)
    cfile = os.path.join(OUT, MODNAME + ".c")
    # language_level is left at the Cython default exactly like setup.py:31 does.
    subprocess.check_call([sys.executable, "-m", "cython", pyx, "-o", cfile],
                          stdout=None if verbose else subprocess.DEVNULL)
    cc = sysconfig.get_config_var("CC") or "gcc"
    cmd = cc.split() + [
        "-shared", "-fPIC", "-O2", "-fwrapv", "-fno-strict-aliasing", "-w",
        "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
        "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
        cfile, "-o", target]
    subprocess.check_call(cmd)
    os.remove(cfile)
    with open(os.path.join(OUT, "__init__.py"), "w") as f:
        f.write("# built by oracle/build_ref.py from the unmodified reference source\n")
    return target


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    p = build(a.reference, a.force)
    print("built" if p else "reference tree absent; nothing built", p or "")
