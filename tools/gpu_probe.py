#!/usr/bin/env python
"""Kernel-level timing sweep on one B200 (development tool; bench.py is the contract).

Times every hot kernel of libcpab_b200 with CUDA events on torch's current stream over the
BASELINE workload shapes and the tuning variants, so that one gpurun call answers "which
configuration, and how far from the roofline".  Writes JSON lines to stdout.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcpab_b200 import Cpab, _lib, ops            # noqa: E402
from libcpab_b200.transformer import _basis         # noqa: E402

# algorithmic flop counts per (point,theta) pair, SURVEY.md 8-d
F_FWD = {1: 400, 2: 1950, 3: 3550}
F_BWD = {1: 1400, 2: 4950, 3: 9550}


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def fma_peak():
    lib = _lib.load()
    out = torch.zeros(1, device="cuda")
    blocks, iters = 148 * 8, 4096
    st = torch.cuda.current_stream().cuda_stream
    med, best = timeit(lambda: _lib.check(lib.cpab_b200_fp32_fma_probe(blocks, iters, out.data_ptr(), st), "probe"))
    flops = blocks * 256 * iters * 64 * 2
    emit(kind="fp32_fma_peak", tflops_med=flops / med / 1e9, tflops_best=flops / best / 1e9)
    return flops / best / 1e9


def run_config(name, tess, n_theta, size, kw, peak, variants=True):
    torch.manual_seed(1234)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    ndim = len(tess)
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    nP = grid.shape[1]
    pairs = n_theta * nP
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    med, best = timeit(lambda: ops.theta_to_trels(theta, Bt, tess, 50))
    emit(kind="theta_to_trels", cfg=name, ms=med, cells_per_s=n_theta * T.params.nC / med * 1e3)
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)

    def fwd_line(tag, **extra):
        for fast in (False, True):
            med, best = timeit(lambda: ops.forward(grid, Tr, tess, 50, fast_math=fast))
            emit(kind="forward", cfg=name, variant=tag, fast_math=fast, ms=med, pairs_per_s=pairs / med * 1e3,
                 tflops_alg=pairs * F_FWD[ndim] / med / 1e9, frac_fp32=pairs * F_FWD[ndim] / med / 1e9 / peak, **extra)

    fwd_line("default")
    _lib.set_tuning("chunk_auto", 1)
    fwd_line("chunk_auto")
    _lib.set_tuning("chunk_auto", 0)
    if variants:
        for ppt in (1, 2):
            for chunk in (1024, 4096):
                _lib.set_tuning("fwd_ppt", ppt)
                _lib.set_tuning("chunk_pts", chunk)
                fwd_line(f"ppt{ppt}_chunk{chunk}")
        _lib.set_tuning("fwd_ppt", 1)
        _lib.set_tuning("chunk_pts", 1024)

    gout = torch.randn(n_theta, ndim, nP, device="cuda")

    def bwd_line(tag):
        med, best = timeit(lambda: ops.backward_theta(grid, As, B, gout, tess, 50))
        emit(kind="backward", cfg=name, variant=tag, ms=med, pairs_per_s=pairs / med * 1e3,
             tflops_alg=pairs * F_BWD[ndim] / med / 1e9, frac_fp32=pairs * F_BWD[ndim] / med / 1e9 / peak)

    bwd_line("default")
    _lib.set_tuning("chunk_auto", 1)
    bwd_line("chunk_auto")
    _lib.set_tuning("chunk_auto", 0)
    for seg, stage, block in ((3, 1, 128), (3, 0, 128), (3, 0, 256), (5, 0, 128), (5, 0, 256)):
        _lib.set_tuning("bwd_seg", seg); _lib.set_tuning("bwd_stage", stage); _lib.set_tuning("bwd_block", block)
        bwd_line(f"seg{seg}_stage{stage}_block{block}")
    _lib.set_tuning("bwd_seg", 0); _lib.set_tuning("bwd_stage", -1); _lib.set_tuning("bwd_block", 128)
    if variants:
        for seg in (5, 10):
            for block in (64, 128, 256):
                for chunk in (1024, 2048):
                    _lib.set_tuning("bwd_seg", seg)
                    _lib.set_tuning("bwd_block", block)
                    _lib.set_tuning("chunk_pts", chunk)
                    bwd_line(f"seg{seg}_block{block}_chunk{chunk}")
        _lib.set_tuning("bwd_seg", 0)
        _lib.set_tuning("bwd_block", 128)
        _lib.set_tuning("chunk_pts", 1024)

    if ndim == 1:   # opt-in closed-form mode (not in the reference)
        med, best = timeit(lambda: ops.forward_closed_form(grid, As, tess))
        emit(kind="forward_closed_form", cfg=name, ms=med, pairs_per_s=pairs / med * 1e3)
        med, best = timeit(lambda: ops.backward_theta_closed_form(grid, As, B, gout, tess))
        emit(kind="backward_closed_form", cfg=name, ms=med, pairs_per_s=pairs / med * 1e3)

    # interpolation on the transformed grid
    C = 1
    data = torch.rand((n_theta, C, *size), device="cuda")
    gt = ops.forward(grid, Tr, tess, 50)
    byts = n_theta * nP * (4 * ndim + 8 * C)
    for var in (0, 1, 2, 3, 4):
        _lib.set_tuning("interp_variant", var)
        med, best = timeit(lambda: ops.interpolate_forward(data, gt, size))
        emit(kind="interp_fwd", cfg=name, variant=var, ms=med, gbps_alg=byts / med / 1e6, points_per_s=pairs / med * 1e3)
    _lib.set_tuning("interp_variant", 2)
    g2 = torch.randn_like(data)
    med, best = timeit(lambda: ops.interpolate_backward(data, gt, g2, True, False))
    byts = n_theta * nP * (8 * ndim + 8 * C)
    emit(kind="interp_bwd_dgrid", cfg=name, ms=med, gbps_alg=byts / med / 1e6)
    med, best = timeit(lambda: ops.interpolate_backward(data, gt, g2, True, True))
    emit(kind="interp_bwd_dgrid_ddata", cfg=name, ms=med)

    # end-to-end through the API (autograd), forward + backward w.r.t. theta
    th = theta.clone().requires_grad_(True)

    def step():
        th.grad = None
        out = T.transform_data(data, th, size)
        (out * g2).sum().backward()

    med, best = timeit(step, warmup=2, iters=3)
    emit(kind="api_fwd_bwd", cfg=name, ms=med, pairs_per_s=pairs / med * 1e3)


def main():
    emit(kind="env", gpu=torch.cuda.get_device_name(0), build=_lib.load().cpab_b200_build_info().decode())
    peak = fma_peak()
    t0 = time.time()
    run_config("cfg1_1d50", [50], 64, [1000], {}, peak, variants=False)
    run_config("cfg2_2d3x3", [3, 3], 64, [256, 256], {}, peak)
    run_config("cfg3_2d10x10vp_b128", [10, 10], 128, [512, 512], {"volume_perservation": True}, peak)
    run_config("cfg4_3d4x4x4_b4", [4, 4, 4], 4, [128, 128, 128], {}, peak)
    run_config("cfg5_1d100_b8192", [100], 8192, [1024], {}, peak)
    emit(kind="done", seconds=time.time() - t0)


if __name__ == "__main__":
    main()
