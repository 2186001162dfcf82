"""Build libtoeplitz_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python toeplitzmatrices.jl_b200/build.py [--force]

Objects go to ``toeplitzmatrices.jl_b200/build/`` and the library to
``toeplitzmatrices.jl_b200/lib/libtoeplitz_b200.so`` (both git-ignored; the ``.so``
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libtoeplitz_b200.so")
SOURCES = ["capi.cu", "comm.cu", "levinson.cu", "levinson_small.cu", "levinson_multiwarp.cu", "levinson_gen.cu", "levinson_gen_mw.cu", "rows_f64.cu", "rows_f32.cu", "cols_f64.cu", "cols_f32.cu", "cols_tma.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_header() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src: str, force: bool, hdr_t: float) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    sp = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(sp), hdr_t)):
        return obj
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc(), *flags, "-c", sp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build_variant(tag: str, defines) -> str:
    """Experiment build: objects and library suffixed with `tag`, extra -D flags (never the default)."""
    objdir = os.path.join(HERE, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + ["-D" + d for d in defines]

    sources = list(SOURCES)
    if any(d.startswith("TB_EXP_CA") for d in defines):
        sources.append(os.path.join("experiments", "cols_ca.cu"))       # fused pass C + pass A launches (option ca_fuse)

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src).replace(".cu", ".o"))
        r = subprocess.run([_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, sources))
    lib = os.path.join(LIBDIR, "libtoeplitz_b200_%s.so" % tag)
    r = subprocess.run([_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_t = _newest_header()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_t), SOURCES))
    if (force or not os.path.exists(LIB)
            or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs)):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(LIB)
    return LIB


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force="--force" in sys.argv, verbose=True)
