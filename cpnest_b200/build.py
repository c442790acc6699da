"""Build libcpnest_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m cpnest_b200.build [--force] [-j N]

The shared library lands in cpnest_b200/lib/ (git-ignored, but it travels to the GPU box with
the gpurun snapshot).  One object per translation unit, compiled in parallel.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libcpnest_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/cpnest_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _unit_digest(src):
    """Digest of one translation unit: its source, every header (all units include most of them) and the flags."""
    h = hashlib.sha256()
    heads = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    for f in [src] + heads + ["../../include/cpnest_b200.h"]:
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    """Compile one unit unless its object is up to date (so a change to api.cu alone recompiles api.cu alone)."""
    out = os.path.join(OBJ, src[:-3] + ".o")
    stamp = out + ".sha256"
    dig = _unit_digest(src)
    if os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return out
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    open(stamp, "w").write(dig + "\n")
    return out


def build(force=False, jobs=None, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(HERE, "lib", "build.sha256")
    dig = _digest()
    if force:
        for f in os.listdir(OBJ):
            if f.endswith(".sha256"):
                os.remove(os.path.join(OBJ, f))
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and no up-to-date prebuilt %s" % (NVCC, LIB))
    srcs = _sources()
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    if verbose:
        print("cpnest_b200.build: compiling %d translation units for sm_100a with %d jobs" % (len(srcs), jobs), flush=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig + "\n")
    if verbose:
        print("cpnest_b200.build: wrote", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-j", type=int, default=None)
    a = ap.parse_args()
    build(force=a.force, jobs=a.j)
