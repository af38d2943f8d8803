"""Development aid: worst trajectories of the n-D hit-time kernel against the checker."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
from conftest import load_golden
from oracle import oracle as O
from libcpab_b200 import ops

for name, npts in [("d2_t3x3", 150), ("d2_t10x10_vp", 150), ("d3_t2x2x2", 120), ("d2_t2x3_free_vp", 80), ("d3_t2x2x2_free", 80)]:
    g = load_golden(name)
    nc = g["nc"].tolist()
    theta = g["theta"][-3:].astype(np.float64)
    As = O.theta_to_affine(g["B"], theta, nc, np.float64)
    grid = g["grid"].astype(np.float64)
    grid = grid[:, ::max(1, grid.shape[1] // npts)]
    ref = O.closed_form_nd(grid, As, nc)
    got = ops.forward_closed_form(torch.from_numpy(grid).cuda(), torch.from_numpy(As).cuda(), nc).cpu().numpy()
    e = np.abs(got - ref).max(axis=1)
    order = np.argsort(e.ravel())[::-1][:6]
    print(name, "max", e.max(), "median", np.median(e), "n>1e-10:", int((e > 1e-10).sum()), "of", e.size)
    for o in order:
        t, i = divmod(int(o), e.shape[1])
        print("   theta", t, "pt", i, grid[:, i], "err", e[t, i], "ref", ref[t, :, i], "got", got[t, :, i])
