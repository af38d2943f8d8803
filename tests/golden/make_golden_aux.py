#!/usr/bin/env python
"""Generate tests/golden/aux_*.npz from the UNMODIFIED reference: `Cpab.calc_vectorfield`
(libcpab/pytorch/functions.py:111-129) and the theta-space covariance that
`Cpab.sample_transformation_with_prior` (libcpab/cpab.py:192-241) builds before sampling.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_aux.py
Same import shims as make_golden.py; the reference runs with backend='pytorch', device='cpu'.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, import_reference  # noqa: E402


def main():
    import torch
    import_reference()
    from libcpab import Cpab
    for name, tess, zb, vp, npts, ls, ov in (("aux_1d_t50", [50], True, False, [300], 0.1, 1.0),
                                             ("aux_2d_t3x3", [3, 3], True, False, [23, 31], 0.1, 1.0),
                                             ("aux_2d_t4x2_vp", [4, 2], False, True, [17, 19], 0.25, 0.7),
                                             ("aux_3d_t2x2x2", [2, 2, 2], True, False, [7, 6, 5], 0.5, 2.0)):
        torch.manual_seed(len(name))
        T = Cpab(tess, backend="pytorch", device="cpu", zero_boundary=zb, volume_perservation=vp, override=True)
        theta = T.sample_transformation(1)
        # strictly interior, off-lattice points (the reference's torch findcellidx mutates its input
        # and disagrees with its own C++ search on/outside the boundary in 3-D)
        gen = torch.Generator().manual_seed(7)
        grid = (0.02 + 0.96 * torch.rand((len(tess), int(np.prod(npts))), generator=gen)).contiguous()
        v = T.calc_vectorfield(grid.clone(), theta)
        # The prior.  As shipped it raises on every backend: cpab.py:218 calls `backend.zeros(D, D,
        # device=...)` but `zeros(*s)` takes no keyword (pytorch/functions.py:64, numpy/functions.py:62),
        # and cpab.py:235 passes the device string as `dtype` to the pytorch `to`.  Its arithmetic
        # (cpab.py:211-237) is backend-independent, so it is run here through the reference's numpy
        # backend with one accommodation -- `zeros` / `ones` take the shape as positional integers and
        # `zeros` ignores `device` -- and the
        # covariance it hands to the sampler is recorded.
        Tn = Cpab(tess, backend="numpy", device="cpu", zero_boundary=zb, volume_perservation=vp, override=False)
        Tn.params.basis = T.params.basis
        rec = {}
        real = Tn.sample_transformation

        def spy(n_sample=1, mean=None, cov=None):
            rec["cov"] = np.array(cov)
            return real(n_sample, mean=mean, cov=cov)

        Tn.sample_transformation = spy
        zeros0, ones0 = Tn.backend.zeros, Tn.backend.ones
        Tn.backend.zeros = lambda *a, device=None: zeros0(a)
        Tn.backend.ones = lambda *a: ones0(a)               # numpy/functions.py:66: np.ones(2, 2) is not a shape either
        try:
            s = Tn.sample_transformation_with_prior(3, length_scale=ls, output_variance=ov)
        finally:
            Tn.backend.zeros, Tn.backend.ones = zeros0, ones0
        np.savez_compressed(os.path.join(OUT, name + ".npz"), B=np.asarray(T.params.basis, dtype=np.float64),
                            nc=np.asarray(tess, dtype=np.int32), zero_boundary=np.bool_(zb),
                            volume_perservation=np.bool_(vp), theta=theta.numpy(), grid=grid.numpy(),
                            vectorfield=v.numpy(), length_scale=np.float64(ls), output_variance=np.float64(ov),
                            cov_theta=rec["cov"], centers=np.asarray(T.tesselation.get_cell_centers()))
        print(name, "v", tuple(v.shape), "cov", tuple(rec["cov"].shape), "sampled" if s is not None else "not sampled")


def aligner_case():
    """`CpabAligner.alignment_by_gradient` (libcpab/alignment.py:60-87): five Adam steps on a smooth
    2-D image pair, by the unmodified reference on the CPU."""
    import torch
    from libcpab import Cpab, CpabAligner
    torch.manual_seed(77)
    T = Cpab([2, 2], backend="pytorch", device="cpu", zero_boundary=True, volume_perservation=False, override=True)
    xs = torch.linspace(0, 1, 28)
    x1 = (torch.sin(5.0 * xs)[:, None] * torch.cos(3.0 * xs)[None, :] + 0.3 * xs[:, None])[None, None].contiguous()
    theta_true = 0.4 * T.sample_transformation(1)
    x2 = T.transform_data(x1, theta_true, outsize=(28, 28)).detach()
    A = CpabAligner(T)
    theta = A.alignment_by_gradient(x1, x2, maxiter=5, lr=1e-2)
    loss_end = float(torch.norm(T.transform_data(x1, theta.detach(), outsize=(28, 28)) - x2))
    np.savez_compressed(os.path.join(OUT, "aux_align_2d.npz"), B=np.asarray(T.params.basis, dtype=np.float64),
                        nc=np.asarray([2, 2], dtype=np.int32), x1=x1.numpy(), x2=x2.numpy(),
                        theta_true=theta_true.numpy(), theta_out=theta.detach().numpy(),
                        loss_start=np.float64(float(torch.norm(x1 - x2))), loss_end=np.float64(loss_end))
    print("aux_align_2d", theta.detach().numpy().ravel()[:4], "loss", float(torch.norm(x1 - x2)), "->", loss_end)


if __name__ == "__main__":
    main()
    aligner_case()
