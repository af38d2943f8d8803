"""The UNMODIFIED reference package, made importable next to this repository.  TEST INFRASTRUCTURE ONLY.

`install()` copies /root/reference/libcpab (where present: the build container) to
baseline/_ref/libcpab -- git-ignored, but it travels to the GPU box with the repository snapshot,
like oracle/_ref.  No file is edited.

`import_with_b200_backend()` imports that package on today's stack (three in-process shims, the same
as tests/golden/make_golden.py: a stub matplotlib, scipy.transpose/compress, torch.solve) with ONE
substitution: `torch.utils.cpp_extension.load(name='cpab_gpu', ...)` -- the call with which
libcpab/pytorch/transformer.py:49-55 JIT-builds the reference's CUDA extension -- returns
integration/cpab_b200.py, the ctypes stub of INTEGRATION.md section 1, instead.  The reference's
own `Cpab(..., backend='pytorch', device='gpu')`, its `_CPABFunction_AnalyticGrad`, its torch expm
and its torch interpolation then run unchanged on top of libcpab_b200.so.
(`cpab_cpu` is not built: the reference then reports `_cpu_succes = False`, which only matters for
CPU tensors.)
"""
import importlib
import os
import shutil
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
REF_SRC = "/root/reference/libcpab"
REF_DST = os.path.join(ROOT, "baseline", "_ref")


def install() -> str | None:
    """Copy the reference package to baseline/_ref/libcpab (no-op when /root/reference is absent)."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    dst = os.path.join(REF_DST, "libcpab")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return REF_DST


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "libcpab", "cpab.py"))


def import_with_b200_backend():
    """Returns the reference's top-level module `libcpab`, bound to libcpab_b200.so for device='gpu'."""
    if "libcpab" in sys.modules and getattr(sys.modules["libcpab"], "_b200_dropin", False):
        return sys.modules["libcpab"]
    if not available():
        raise RuntimeError("baseline/_ref/libcpab is missing (run __graft_entry__.build() where /root/reference exists)")
    import numpy as np
    import scipy
    import torch
    import torch.utils.cpp_extension as cpp_ext

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.figure = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    scipy.transpose = np.transpose
    scipy.compress = np.compress
    # torch.solve survives only as a stub that raises (pytorch/expm.py:27 calls it); solve(B, A) solved A X = B
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    if not hasattr(np, "bool"):
        np.bool = bool

    from libcpab_b200 import _lib
    _lib.load()                                                     # builds / checks the library
    os.environ["LIBCPAB_B200_SO"] = _lib.library_path()
    spec = importlib.util.spec_from_file_location("cpab_b200_stub", os.path.join(ROOT, "integration", "cpab_b200.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)

    real_load = cpp_ext.load

    def load(name, sources=None, **kw):
        if name == "cpab_gpu":
            return stub
        raise RuntimeError("%s is not built in this harness" % name)

    cpp_ext.load = load
    sys.path.insert(0, REF_DST)
    try:
        ref = importlib.import_module("libcpab")
        # (the backend is imported lazily by Cpab.__init__: pull it in while `load` is substituted)
        rt = importlib.import_module("libcpab.pytorch.transformer")
        importlib.import_module("libcpab.pytorch.functions")
    finally:
        cpp_ext.load = real_load
        sys.path.remove(REF_DST)
    assert rt._gpu_succes and rt.cpab_gpu is stub
    ref._b200_dropin = True
    return ref
