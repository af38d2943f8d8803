#!/usr/bin/env python
"""Development: backward with the 64-register build (LIBCPAB_B200_SO) and segment lengths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libcpab_b200 import Cpab, _lib, ops
from libcpab_b200.transformer import _basis
from tools.gpu_probe import timeit, emit, F_BWD
def run(name, tess, n_theta, size, kw):
    torch.manual_seed(1234)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    theta = T.sample_transformation(n_theta); grid = T.uniform_meshgrid(size)
    nP = grid.shape[1]; B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)
    gout = torch.randn(n_theta, len(tess), nP, device="cuda")
    for seg in (3, 5):
        _lib.set_tuning("bwd_seg", seg)
        med, best = timeit(lambda: ops.backward_theta(grid, As, B, gout, tess, 50))
        emit(kind="backward", lib=os.environ.get("LIBCPAB_B200_SO", "default"), cfg=name, seg=seg, ms=med)
run("cfg2_2d3x3", [3, 3], 64, [256, 256], {})
run("cfg3_2d10x10vp_b128", [10, 10], 128, [512, 512], {"volume_perservation": True})
