# libcpab/pytorch/cpab_b200.py -- ctypes stub with the interface of the reference's JIT module `cpab_gpu`
# (pybind11 entry points `forward` / `backward`, libcpab/pytorch/transformer_cuda.cpp:17-76).
#
# This is the file INTEGRATION.md section 1 asks a maintainer of the reference to add: with
#     from . import cpab_b200 as cpab_gpu; _gpu_succes = True
# in place of the `load(name='cpab_gpu', ...)` block (libcpab/pytorch/transformer.py:49-69) the
# reference's own Cpab / _CPABFunction_AnalyticGrad run unchanged on libcpab_b200.so.
# tests/test_gpu_reference_dropin.py does exactly that with the unmodified reference package.
import ctypes
import os

import torch

_lib = ctypes.CDLL(os.environ.get("LIBCPAB_B200_SO", "libcpab_b200.so"))
_vp, _i, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
_ip = ctypes.POINTER(ctypes.c_int)
_lib.cpab_b200_forward.argtypes = [_i, _i, _i, _ip, _i, _i, _l, _i, _vp, _vp, _vp, _vp]
_lib.cpab_b200_backward_jacobian.argtypes = [_i, _i, _ip, _i, _i, _i, _l, _i, _vp, _vp, _vp, _vp, _vp]
_lib.cpab_b200_last_error.restype = ctypes.c_char_p


def _check(rc):
    if rc != 0:
        raise RuntimeError(_lib.cpab_b200_last_error().decode())


def _require(t, name):                            # transformer_cuda.cpp:12-14 (CHECK_INPUT)
    if not t.is_cuda:
        raise RuntimeError(name + " must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(name + " must be contiguous")


def _geom(points, n_theta):                       # transformer_cuda.cpp:28 (broadcast rule)
    bc = int(points.dim() == 3 and points.size(0) == n_theta)
    return bc, (points.size(1) if bc else points.size(0)), (points.size(2) if bc else points.size(1))


def forward(points, trels, nstepsolver, nc):      # replaces transformer_cuda.cpp:17-41
    _require(points, "points"); _require(trels, "trels")
    bc, ndim, nP = _geom(points, trels.size(0))
    out = torch.empty(trels.size(0), ndim, nP, device=points.device)
    ncs = (ctypes.c_int * ndim)(*nc.tolist())
    with torch.cuda.device(points.device):
        _check(_lib.cpab_b200_forward(0, 0, ndim, ncs, int(nstepsolver), trels.size(0), nP, bc,
                                      points.data_ptr(), trels.data_ptr(), out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    return out


def backward(points, As, Bs, nstepsolver, nc):    # replaces transformer_cuda.cpp:43-70
    _require(points, "points"); _require(As, "As"); _require(Bs, "Bs")
    bc, ndim, nP = _geom(points, As.size(0))
    d, n_theta = Bs.size(0), As.size(0)
    jac = torch.empty(d, n_theta, ndim, nP, device=points.device)
    ncs = (ctypes.c_int * ndim)(*nc.tolist())
    with torch.cuda.device(points.device):
        _check(_lib.cpab_b200_backward_jacobian(0, ndim, ncs, int(nstepsolver), n_theta, d, nP, bc,
                                                points.data_ptr(), As.data_ptr(), Bs.data_ptr(),
                                                jac.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return jac
