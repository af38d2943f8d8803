#!/usr/bin/env python
"""Small driver for `ncu`: launches each hot kernel a few times on a BASELINE-shaped problem.

    ncu --set full --clock-control none --import-source on -k regex:'k_forward|k_backward|k_interp' \
        -o gpurun_out/prof python tools/ncu_target.py [2d|3d|1d]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcpab_b200 import Cpab, ops                      # noqa: E402
from libcpab_b200.transformer import _basis             # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "2d"
tess, n_theta, size, kw = {
    "2d": ([10, 10], 64, [512, 512], {"volume_perservation": True}),
    "3d": ([4, 4, 4], 2, [128, 128, 128], {}),
    "1d": ([100], 8192, [1024], {}),
}[which]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
T = Cpab(tess, backend="pytorch", device="gpu", **kw)
theta = T.sample_transformation(n_theta)
grid = T.uniform_meshgrid(size)
B, Bt = _basis(T.params, theta.device, theta.dtype)
data = torch.rand(n_theta, 1, *size, device="cuda")
gout = torch.randn(n_theta, len(tess), grid.shape[1], device="cuda")
g2 = torch.randn_like(data)
for _ in range(reps):
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)
    if which == "1d" or "closed" in sys.argv:
        x1 = ops.forward_closed_form(grid, As, tess)
        ops.backward_theta_closed_form(grid, As, B, gout, tess, newpoints=x1 if which != "1d" else None)   # as autograd calls it
    gt = ops.forward(grid, Tr, tess, 50)
    out = ops.interpolate_forward(data, gt, size)
    ops.interpolate_backward(data, gt, g2, True, False)
    ops.backward_theta(grid, As, B, gout, tess, 50)
    ops.backward_theta(grid, As, B, gout, tess, 50, fast_grad=True)
torch.cuda.synchronize()
print("done", which)
