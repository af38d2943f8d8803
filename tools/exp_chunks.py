#!/usr/bin/env python
"""Development sweep: chunk size / keep variants of k_forward and k_backward (one B200)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcpab_b200 import Cpab, _lib, ops
from libcpab_b200.transformer import _basis
from tools.gpu_probe import timeit, emit, F_FWD, F_BWD, fma_peak

def run(name, tess, n_theta, size, kw, peak):
    torch.manual_seed(1234)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    ndim = len(tess)
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    nP = grid.shape[1]; pairs = n_theta * nP
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)
    gout = torch.randn(n_theta, ndim, nP, device="cuda")
    for chunk in (0, 512, 1024, 2048, 4096):
        if chunk: _lib.set_tuning("chunk_pts", chunk)
        else: _lib.set_tuning("chunk_auto", 1)
        med, best = timeit(lambda: ops.forward(grid, Tr, tess, 50))
        emit(kind="forward", cfg=name, chunk=chunk, ms=med, frac=pairs * F_FWD[ndim] / med / 1e9 / peak)
        med, best = timeit(lambda: ops.backward_theta(grid, As, B, gout, tess, 50))
        emit(kind="backward", cfg=name, chunk=chunk, ms=med, frac=pairs * F_BWD[ndim] / med / 1e9 / peak)
    _lib.set_tuning("chunk_pts", 1024); _lib.set_tuning("chunk_auto", 1)

peak = fma_peak()
run("cfg2_2d3x3", [3, 3], 64, [256, 256], {}, peak)
run("cfg3_2d10x10vp_b128", [10, 10], 128, [512, 512], {"volume_perservation": True}, peak)
run("cfg4_3d4x4x4_b4", [4, 4, 4], 4, [128, 128, 128], {}, peak)
run("cfg5_1d100_b8192", [100], 8192, [1024], {}, peak)
run("cfg1_1d50", [50], 64, [1000], {}, peak)
