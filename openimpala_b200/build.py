"""In-tree build of the CUDA library, the host C++ front end and the oracle.

Everything is compiled with explicit nvcc / g++ / gcc command lines (no build
system needed on the GPU box: the built files travel with the tree).

    python -m openimpala_b200.build            # everything
    python -m openimpala_b200.build lib        # CUDA shared library only
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "openimpala_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIBDIR = os.path.join(PKG, "lib")
BINDIR = os.path.join(PKG, "bin")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libopenimpala_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"] + ARCH

CU_SOURCES = ["oi_level0.cu", "oi_level0_ring.cu", "oi_level0_tma.cu", "oi_level0_pair.cu", "oi_coarse.cu", "oi_vecops.cu", "oi_mask.cu", "oi_halo.cu", "oi_refabi.cu", "oi_solver.cu"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build_lib(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "openimpala_b200.h"))
    objs = []
    jobs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            jobs.append((s, o))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(lambda so: _run([NVCC] + NVCC_FLAGS + ["-c", so[0], "-o", so[1]],
                                    log=so[1] + ".log"), jobs))
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static", "-ldl"])
    return LIB


def build_host(force: bool = False):
    """Host C++ front end (AMReX-shaped shims, TortuosityHypre / VolumeFraction
    classes, readers, the Diffusion app and the reference-style test drivers)."""
    if not os.path.isdir(HOST):
        return []
    os.makedirs(BINDIR, exist_ok=True)
    mk = os.path.join(HOST, "Makefile")
    if os.path.exists(mk):
        _run(["make", "-s", "-C", HOST, "-j", str(os.cpu_count() or 1)] + (["-B"] if force else []))
    return [os.path.join(BINDIR, f) for f in os.listdir(BINDIR)]


def build_oracle(force: bool = False):
    odir = os.path.join(ROOT, "oracle")
    mk = os.path.join(odir, "Makefile")
    if os.path.exists(mk):
        _run(["make", "-s", "-C", odir] + (["-B"] if force else []))


def build_all(force: bool = False):
    build_lib(force)
    build_host(force)
    build_oracle(force)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    force = "--force" in sys.argv
    if what == "lib":
        print(build_lib(force))
    elif what == "host":
        print(build_host(force))
    elif what == "oracle":
        build_oracle(force)
    elif what == "clean":
        shutil.rmtree(OBJDIR, ignore_errors=True)
    else:
        build_all(force)
        print("built", LIB)
