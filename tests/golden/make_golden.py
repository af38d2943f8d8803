#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (SkafteNicki/libcpab).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does
------------
* copies ``/root/reference/libcpab`` to a scratch directory (the reference writes its basis
  pickles next to the package, and /root/reference is read-only) -- no file is edited;
* installs three in-process shims so the 2019 host code imports under today's stack
  (SURVEY.md 8-c): a stub ``matplotlib``; ``scipy.transpose/compress`` = numpy's;
  ``torch.solve(B, A)`` = ``torch.linalg.solve(A, B)``;
* lets the reference JIT-build its own CPU extension (``cpab_cpu``: pytorch/transformer.cpp +
  core/cpab_ops.cpp) and drives everything through the reference's public ``Cpab`` API with
  ``backend='pytorch', device='cpu'`` (plus the numpy backend for the float64 forward);
* stores inputs (basis B, theta, grids, data, upstream gradients) and the reference's outputs.

The fixtures pin the oracle (tests/test_oracle_pinned.py) and are the committed known answers the
GPU parity tests compare against (tests/test_gpu_golden.py).  Nothing in tests/ reads
/root/reference at run time.
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    scratch = tempfile.mkdtemp(prefix="libcpab_ref_")
    shutil.copytree(os.path.join(REF, "libcpab"), os.path.join(scratch, "libcpab"))
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(scratch, "torch_ext"))

    # shim 1: matplotlib is only used for plotting helpers (cpab.py:11, default args :349)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.figure = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    # shim 2: names removed from scipy's top level (core/utility.py:17)
    import scipy
    scipy.transpose = np.transpose
    scipy.compress = np.compress
    # shim 3: torch.solve removed in torch 2 (pytorch/expm.py:27); solve(B, A) solved A X = B
    import torch
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    # np.bool was removed from numpy (core/tesselation.py:243, only hit when zero_boundary=False)
    if not hasattr(np, "bool"):
        np.bool = bool

    sys.path.insert(0, scratch)
    import libcpab  # noqa: F401  (the reference)
    return scratch


def cell_probe_points(ndim, n, rng, nc):
    """Random points in and slightly outside the unit box plus points on cell faces/diagonals."""
    pts = rng.uniform(-0.15, 1.15, size=(ndim, n)).astype(np.float32)
    inside = rng.uniform(0.0, 1.0, size=(ndim, n)).astype(np.float32)
    # lattice points: multiples of 1/(2*nc) hit vertices, edge mid-points, centres, diagonals
    lat = np.stack([rng.integers(0, 2 * nc[j] + 1, size=n) / (2.0 * nc[j]) for j in range(ndim)])
    lat = lat.astype(np.float32)
    # ulp neighbours of lattice points
    up = np.nextafter(lat, np.float32(2.0))
    dn = np.nextafter(lat, np.float32(-1.0))
    # points exactly on the unit-box faces
    face = inside.copy()
    sel = rng.integers(0, ndim, size=n)
    face[sel, np.arange(n)] = rng.integers(0, 2, size=n).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([pts, inside, lat, up, dn, face], axis=1))


def main():
    import torch
    import_reference()
    from libcpab import Cpab, CpabSequential
    from libcpab.pytorch import transformer as rtrans
    from libcpab.pytorch.expm import expm as ref_expm

    assert rtrans._cpu_succes, "reference CPU extension failed to JIT-build"
    cpab_cpu = rtrans.cpab_cpu
    torch.set_num_threads(1)

    # ---------------------------------------------------------------- cell indices (C++ core)
    # The extension does not export findcellidx; forward() with ONE step and Trels whose only
    # non-zero entry is translation_0 = cell index returns exactly that index per point.
    cells = {}
    rng = np.random.default_rng(20261017)
    for nc in ([1], [7], [50], [100], [1, 1], [3, 3], [10, 10], [5, 2], [2, 7], [1, 1, 1],
               [2, 2, 2], [4, 4, 4], [3, 2, 5]):
        ndim = len(nc)
        nC = int({1: 1, 2: 4, 3: 5}[ndim] * np.prod(nc))
        pts = cell_probe_points(ndim, 4000, rng, nc)
        T = np.zeros((1, nC, ndim, ndim + 1), dtype=np.float32)
        T[0, :, 0, ndim] = np.arange(nC, dtype=np.float32)
        out = cpab_cpu.forward(torch.from_numpy(pts), torch.from_numpy(T),
                               torch.tensor(1, dtype=torch.int32),
                               torch.tensor(nc, dtype=torch.int32)).numpy()
        key = "x".join(map(str, nc))
        cells["pts_" + key] = pts
        cells["idx_" + key] = out[0, 0].astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "cells.npz"), **cells)
    print("cells.npz", {k: v.shape for k, v in cells.items() if k.startswith("idx")})

    # ---------------------------------------------------------------- expm (torch Pade-13)
    rng = np.random.default_rng(7)
    ex = {}
    for m in (2, 3, 4):
        A = rng.normal(size=(64, m, m)).astype(np.float32)
        A[:, m - 1, :] = 0
        A[:8] *= 0.02
        A[8:16] *= 8.0        # forces squarings
        A[16] = 0
        ex[f"A{m}"] = A
        ex[f"E{m}"] = ref_expm(torch.from_numpy(A)).numpy()
        ex[f"E{m}_f64"] = ref_expm(torch.from_numpy(A.astype(np.float64))).numpy()
    np.savez_compressed(os.path.join(OUT, "expm.npz"), **ex)

    # ---------------------------------------------------------------- op + API level
    def run_case(name, tess, zb, vp, n_theta, grid_n, data_shape, seed, theta_scale=1.0,
                 outside=False, with_numpy=True):
        torch.manual_seed(seed)
        T = Cpab(tess, backend="pytorch", device="cpu", zero_boundary=zb,
                 volume_perservation=vp, override=True)
        p = T.params
        ndim = p.ndim
        theta = (theta_scale * T.sample_transformation(n_theta)).contiguous()
        grid = T.uniform_meshgrid(grid_n)
        if outside:  # stretch a copy of the grid beyond the unit box (valid-outside mode)
            grid = (grid * 1.3 - 0.15).contiguous()
        rec = dict(B=np.asarray(p.basis, dtype=np.float64), theta=theta.numpy(),
                   grid=grid.numpy(), nc=np.asarray(tess, dtype=np.int32),
                   nstepsolver=np.int32(p.nstepsolver), zero_boundary=np.bool_(zb),
                   volume_perservation=np.bool_(vp), grid_n=np.asarray(grid_n, dtype=np.int32))

        # host pieces exactly as _CPABFunction_AnalyticGrad.forward computes them
        Bt = torch.Tensor(p.basis)
        As = torch.matmul(Bt, theta.t()).t().reshape(n_theta * p.nC, *p.Ashape)
        sq = torch.cat([As, torch.zeros(n_theta * p.nC, 1, ndim + 1)], dim=1)
        Trels = ref_expm((1.0 / p.nstepsolver) * sq)[:, :ndim, :].reshape(n_theta, p.nC, *p.Ashape)
        rec["As"] = As.reshape(n_theta, p.nC, *p.Ashape).numpy()
        rec["Trels"] = Trels.contiguous().numpy()

        # transform_grid forward + backward through the reference API
        th = theta.clone().requires_grad_(True)
        gt = T.transform_grid(grid, th)
        gen = torch.Generator().manual_seed(seed + 3087)
        gout = torch.randn(gt.shape, generator=gen)
        (gt * gout).sum().backward()
        rec["grid_t"] = gt.detach().numpy()
        rec["gout"] = gout.numpy()
        rec["dtheta"] = th.grad.numpy()

        # raw Jacobian straight from the extension (small cases only)
        Bs = Bt.t().reshape(-1, p.nC, *p.Ashape).contiguous()
        if n_theta * p.d * grid.shape[1] * ndim <= 500_000:
            jac = cpab_cpu.backward(grid.contiguous(), As.reshape(n_theta, p.nC, *p.Ashape)
                                    .contiguous(), Bs, torch.tensor(p.nstepsolver,
                                    dtype=torch.int32), torch.tensor(tess, dtype=torch.int32))
            rec["jac"] = jac.numpy()

        # float64 forward by the reference's numpy backend (scipy expm), same basis
        if with_numpy:
            Tn = Cpab(tess, backend="numpy", device="cpu", zero_boundary=zb,
                      volume_perservation=vp, override=False)
            Tn.params.basis = p.basis
            rec["grid_t_numpy64"] = Tn.transform_grid(grid.numpy().astype(np.float64),
                                                      theta.numpy().astype(np.float64))

        # interpolate / transform_data forward + backward (theta and data)
        if data_shape is not None:
            gen = torch.Generator().manual_seed(seed + 11)
            data = torch.rand((n_theta,) + tuple(data_shape), generator=gen)
            th = theta.clone().requires_grad_(True)
            dd = data.clone().requires_grad_(True)
            out = T.transform_data(dd, th, outsize=grid_n)
            r = torch.randn(out.shape, generator=gen)
            (out * r).sum().backward()
            rec["data"] = data.numpy()
            rec["data_t"] = out.detach().numpy()
            rec["data_gout"] = r.numpy()
            rec["data_dtheta"] = th.grad.numpy()
            rec["data_ddata"] = dd.grad.numpy()
            # interpolate alone on the (fixed) transformed grid, with d/dgrid
            g2 = gt.detach().clone().requires_grad_(True)
            o2 = T.interpolate(data, g2, grid_n)
            (o2 * r).sum().backward()
            rec["interp_out"] = o2.detach().numpy()
            rec["interp_dgrid"] = g2.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "d=%d nC=%d" % (p.d, p.nC), "max|dtheta|=%.3g" % np.abs(rec["dtheta"]).max())

    # BASELINE configs[0] at full size
    run_case("cfg1_1d50", [50], True, False, 64, [1000], None, seed=1235)
    # 1-D with data (configs[4] family), valid outside, volume preserving variants
    run_case("d1_t100", [100], True, False, 6, [257], (2, 300), seed=11)
    run_case("d1_t10_free", [10], False, False, 5, [200], (1, 64), seed=12, outside=True)
    # 2-D: configs[1] family (tess [3,3]) and configs[2] family ([10,10]+vp) at reduced size
    run_case("d2_t3x3", [3, 3], True, False, 4, [32, 32], (2, 24, 20), seed=1236)
    run_case("d2_t10x10_vp", [10, 10], True, True, 3, [40, 24], (1, 33, 47), seed=1237)
    run_case("d2_t2x3_free_vp", [2, 3], False, True, 3, [20, 20], (1, 16, 16), seed=13,
             outside=True)
    # 3-D: configs[3] family at reduced size (grid faces hit the coord==1.0 quirk)
    run_case("d3_t2x2x2", [2, 2, 2], True, False, 2, [9, 8, 7], (1, 6, 7, 8), seed=1238)
    run_case("d3_t2x2x2_free", [2, 2, 2], False, False, 2, [6, 6, 6], (2, 5, 5, 5), seed=14,
             outside=True, theta_scale=0.5)
    run_case("d3_t3x2x2_vp", [3, 2, 2], True, True, 2, [8, 8, 8], None, seed=15)

    # ---------------------------------------------------------------- CpabSequential (configs[4])
    torch.manual_seed(1239)
    Ts = [Cpab([20], backend="pytorch", device="cpu", zero_boundary=True, override=True)
          for _ in range(3)]
    S = CpabSequential(*Ts)
    thetas = [t.sample_transformation(5).clone().requires_grad_(True) for t in Ts]
    gen = torch.Generator().manual_seed(99)
    data = torch.rand((5, 2, 96), generator=gen)
    out = S.transform_data(data, thetas, outsize=[128])
    r = torch.randn(out.shape, generator=gen)
    (out * r).sum().backward()
    rec = dict(B=np.asarray(Ts[0].params.basis), nc=np.asarray([20], dtype=np.int32),
               data=data.numpy(), data_t=out.detach().numpy(), data_gout=r.numpy(),
               grids=np.stack([g.detach().numpy() for g in
                               S.transform_grid(S.uniform_meshgrid([128]), thetas,
                                                output_all=True)]))
    for i, t in enumerate(thetas):
        rec[f"theta{i}"] = t.detach().numpy()
        # the reference returns no gradient for `points`, so only the last warp gets one
        rec[f"dtheta{i}"] = (t.grad.numpy() if t.grad is not None
                             else np.zeros_like(t.detach().numpy()))
        rec[f"has_grad{i}"] = np.bool_(t.grad is not None)
    np.savez_compressed(os.path.join(OUT, "seq_1d20x3.npz"), **rec)
    print("seq_1d20x3", [bool(rec[f"has_grad{i}"]) for i in range(3)])


if __name__ == "__main__":
    main()
