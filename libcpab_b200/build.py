"""Build libcpab_b200.so in-tree with nvcc for sm_100a.

    python -m libcpab_b200.build          # or: from libcpab_b200.build import build; build()

The shared object is written next to the sources (libcpab_b200/csrc/libcpab_b200.so); it is
git-ignored (history stays source-only) but travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["cpab_abi.cu", "cpab_integrate.cu", "cpab_adjoint_1d.cu", "cpab_adjoint_2d.cu", "cpab_adjoint_3d.cu",
           "cpab_expm.cu", "cpab_interp.cu", "cpab_probe.cu", "cpab_closed1d.cu", "cpab_closednd.cu"]
HEADERS = ["cpab_common.cuh", "cpab_cell.cuh", "cpab_f32x2.cuh", "cpab_sample.cuh", "cpab_device.cuh", "cpab_adjoint.cuh",
           os.path.join("..", "..", "include", "libcpab_b200.h")]
LIB = os.path.join(CSRC, "libcpab_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libcpab_b200 has no non-CUDA build")
    return nvcc


STAMP = LIB + ".srchash"


def _source_hash() -> str:
    """Content hash of every source and header (mtimes do not survive a repository snapshot)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    """True when the library is absent or was built from different sources."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
