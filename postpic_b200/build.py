"""Build libptg_b200.so (sm_100a only) in-tree with nvcc.

    python -m postpic_b200.build [--force] [--verbose]

The library lands in postpic_b200/lib/ (git-ignored, shipped to the GPU box by
gpurun).  nvcc cross-compiles without a GPU.  cudart is linked statically, so the
library has no run-time dependency besides libcuda (resolved lazily on first use).
"""
import argparse
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libptg_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

SOURCES = ["ptg_capi.cu", "ptg_deposit_global.cu", "ptg_deposit_tiled.cu", "ptg_deposit_tiled_nd1.cu", "ptg_deposit_tiled_nd2.cu", "ptg_deposit_tiled_nd3.cu",
           "ptg_deposit_sorted.cu", "ptg_deposit_runs.cu", "ptg_deposit_runs_nd1.cu", "ptg_deposit_runs_nd2.cu", "ptg_deposit_runs_nd3.cu", "ptg_deposit_runs_bulkwide.cu", "ptg_deposit_fixed.cu", "ptg_util.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # belt and braces: the stencil code uses explicit _rn intrinsics anyway
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(INCLUDE, "ptg_b200.h"))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
