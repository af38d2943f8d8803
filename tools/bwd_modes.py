#!/usr/bin/env python
"""Time cpab_b200_backward_theta in its default (certified) and fast_grad modes on the BASELINE
shapes (development tool).  usage: python tools/bwd_modes.py [cfg ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcpab_b200 import Cpab, _lib, ops               # noqa: E402
from libcpab_b200.transformer import _basis            # noqa: E402

CFGS = {
    "cfg1": ([50], 64, [1000], {}),
    "cfg2": ([3, 3], 64, [256, 256], {}),
    "cfg3": ([10, 10], 128, [512, 512], {"volume_perservation": True}),
    "cfg4": ([4, 4, 4], 4, [128, 128, 128], {}),
    "cfg5": ([100], 8192, [1024], {}),
}
F_BWD = {1: 1400, 2: 4950, 3: 9550}
F_FWD = {1: 400, 2: 1950, 3: 3550}


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name in (sys.argv[1:] or list(CFGS)):
    tess, n_theta, size, kw = CFGS[name]
    torch.manual_seed(1)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)
    nP = grid.shape[1]
    gout = torch.randn(n_theta, len(tess), nP, device="cuda")
    pairs = n_theta * nP
    rec = {"cfg": name, "pairs": pairs}
    rec["fwd_ms"] = timeit(lambda: ops.forward(grid, Tr, tess, 50))
    for fast in (False, True):
        ms = timeit(lambda: ops.backward_theta(grid, As, B, gout, tess, 50, fast_grad=fast))
        rec["bwd_fast_ms" if fast else "bwd_default_ms"] = ms
    for fast in (False, True):          # kernel-level split (events on the launch stream, inside the library)
        _lib.profile_enable(True)
        for _ in range(3):
            ops.backward_theta(grid, As, B, gout, tess, 50, fast_grad=fast)
        torch.cuda.synchronize()
        for slot in ("backward", "backward_redo", "epilogue"):
            ms, n = _lib.profile_read(slot)
            if n:
                rec["k_%s_%s_ms" % (slot, "fast" if fast else "default")] = round(ms / n, 4)
        _lib.profile_enable(False)
    rec["bwd_default_tflops"] = pairs * F_BWD[len(tess)] / rec["bwd_default_ms"] / 1e9
    rec["bwd_fast_tflops"] = pairs * F_BWD[len(tess)] / rec["bwd_fast_ms"] / 1e9
    rec["fwd_tflops"] = pairs * F_FWD[len(tess)] / rec["fwd_ms"] / 1e9
    print(json.dumps(rec), flush=True)
