"""The reference's own CUDA kernels as an on-GPU comparator -- TEST / BENCH INFRASTRUCTURE ONLY.

``oracle/_ref/libcpab_ref_cuda.so`` is the reference's ``libcpab/core/cpab_ops.cu`` compiled
unmodified for sm_100a (``oracle/Makefile``, only where ``/root/reference`` exists; the built
library travels with the repository snapshot).  Only ``bench.py`` and ``tests/`` may import this
module; the product package ``libcpab_b200`` never does.  The wrappers take CUDA tensors and launch
on torch's current stream with the launch configurations of
``libcpab/pytorch/transformer_cuda.cu:18-119``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcpab_ref_cuda.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        vp, i = ctypes.c_void_p, ctypes.c_int
        _lib.cpab_refcuda_forward.restype = i
        _lib.cpab_refcuda_forward.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp]
        _lib.cpab_refcuda_backward.restype = i
        _lib.cpab_refcuda_backward.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    return _lib


def _consts(nc, nsteps, device):
    import torch
    return (torch.tensor([int(nsteps)], dtype=torch.int32, device=device),
            torch.tensor([int(v) for v in nc], dtype=torch.int32, device=device))


def forward(points, trels, nc, nsteps, out):
    """cpab_gpu.forward of the reference: out [n_theta, ndim, nP] (float32 CUDA tensors)."""
    import torch
    n_theta = trels.shape[0]
    broadcast = int(points.dim() == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    ns, ncd = _consts(nc, nsteps, points.device)
    rc = lib().cpab_refcuda_forward(out.data_ptr(), points.data_ptr(), trels.data_ptr(), ns.data_ptr(),
                                    ncd.data_ptr(), int(ndim), int(nP), int(n_theta), broadcast,
                                    torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"reference CUDA forward launch failed: cudaError {rc}")
    return out


def backward(points, As, Bs, nc, nsteps, grad):
    """cpab_gpu.backward of the reference: grad [d, n_theta, ndim, nP], zeroed by the caller."""
    import torch
    n_theta, d, nC = As.shape[0], Bs.shape[0], Bs.shape[1]
    broadcast = int(points.dim() == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    ns, ncd = _consts(nc, nsteps, points.device)
    rc = lib().cpab_refcuda_backward(grad.data_ptr(), points.data_ptr(), As.data_ptr(), Bs.data_ptr(),
                                     ns.data_ptr(), ncd.data_ptr(), int(ndim), int(nP), int(n_theta),
                                     int(d), int(nC), broadcast, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"reference CUDA backward launch failed: cudaError {rc}")
    return grad
