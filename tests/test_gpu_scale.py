"""Size-independent properties at (or near) BASELINE.json's full sizes, where the CPU oracle
would take hours: identity, sub-sample agreement with the oracle, linearity of the gradient in
the upstream gradient, agreement between tiling variants, sharded == unsharded.
"""
import numpy as np
import pytest
import torch

from conftest import assert_grad_parity, rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _T(tess, **kw):
    from libcpab_b200 import Cpab
    return Cpab(tess, backend="pytorch", device="gpu", **kw)


@pytest.mark.parametrize("tess,n_theta,size,kw", [
    ([50], 64, [1000], {}),
    ([3, 3], 64, [256, 256], {}),
    ([10, 10], 32, [512, 512], {"volume_perservation": True}),
    ([4, 4, 4], 2, [128, 128, 128], {}),
    ([100], 8192, [1024], {}),
])
def test_full_size_forward_subsample_against_oracle(tess, n_theta, size, kw):
    from libcpab_b200 import ops
    from libcpab_b200.transformer import _basis
    torch.manual_seed(7)
    T = _T(tess, **kw)
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    out = T.transform_grid(grid, theta)
    assert tuple(out.shape) == (n_theta, len(tess), int(np.prod(size)))
    assert bool(torch.isfinite(out).all())
    # identity: theta = 0 leaves every point where it is, exactly
    same = T.transform_grid(grid, T.identity(2))
    assert bool((same == grid[None]).all())
    # oracle on a sub-sample (same Trels => bit-identical expected)
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, Tr = ops.theta_to_trels(theta[:2], Bt, tess, 50)
    rng = np.random.default_rng(3)
    sel = np.sort(rng.choice(grid.shape[1], size=min(4096, grid.shape[1]), replace=False))
    ref = O.forward(grid[:, sel].cpu().numpy(), Tr.cpu().numpy(), tess, 50)
    assert np.array_equal(out[:2][:, :, sel].cpu().numpy(), ref)


@pytest.mark.parametrize("tess,n_theta,size,kw", [
    ([3, 3], 16, [256, 256], {}),
    ([10, 10], 8, [512, 512], {"volume_perservation": True}),
    ([4, 4, 4], 1, [64, 64, 64], {}),
    ([100], 4096, [1024], {}),
])
def test_full_size_gradient_properties(tess, n_theta, size, kw):
    from libcpab_b200 import _lib, ops
    from libcpab_b200.transformer import _basis
    torch.manual_seed(11)
    T = _T(tess, **kw)
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, _ = ops.theta_to_trels(theta, Bt, tess, 50)
    nP = grid.shape[1]
    g1 = torch.randn(n_theta, len(tess), nP, device="cuda")
    g2 = torch.randn(n_theta, len(tess), nP, device="cuda")
    d1, _ = ops.backward_theta(grid, As, B, g1, tess, 50)
    d2, _ = ops.backward_theta(grid, As, B, g2, tess, 50)
    d12, _ = ops.backward_theta(grid, As, B, 2.0 * g1 - 0.5 * g2, tess, 50)
    # linear in the upstream gradient
    assert rel_err((2.0 * d1 - 0.5 * d2).cpu().numpy(), d12.cpu().numpy()) < 2e-5
    # additive over a split of the point set (what point-sharding relies on)
    half = nP // 2
    da, _ = ops.backward_theta(grid[:, :half].contiguous(), As, B, g1[:, :, :half].contiguous(), tess, 50)
    db, _ = ops.backward_theta(grid[:, half:].contiguous(), As, B, g1[:, :, half:].contiguous(), tess, 50)
    assert rel_err((da + db).cpu().numpy(), d1.cpu().numpy()) < 2e-5
    # tiling variants agree
    try:
        _lib.set_tuning("bwd_seg", 10)
        _lib.set_tuning("chunk_pts", 512)
        dv, _ = ops.backward_theta(grid, As, B, g1, tess, 50)
    finally:
        _lib.set_tuning("bwd_seg", 0)
        _lib.set_tuning("chunk_pts", 1024)
        _lib.set_tuning("chunk_auto", 1)
    assert rel_err(dv.cpu().numpy(), d1.cpu().numpy()) < 2e-5
    # oracle (the reference's float32 arithmetic) on a sub-sample of points, first thetas.
    # The discretised flow is piecewise affine in p, so its Jacobian jumps across cell faces: a
    # point whose float32 RK2 iterate lands within an ulp of a face can be assigned to the
    # neighbouring cell by any implementation that does not reproduce the reference's roundings bit
    # for bit (its own CUDA build, FMA-contracted by nvcc, included).  Such a flip happens about
    # once per 1e6 (point, step) events and moves that theta's gradient by ~h |dA| / nP ~ 1e-4
    # relative; everything else agrees to rounding (conftest.assert_grad_parity).
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(nP, size=min(512, nP), replace=False))
    Bs = np.ascontiguousarray(B.cpu().numpy().T.reshape(B.shape[1], -1, len(tess), len(tess) + 1))
    n_chk = min(8, n_theta)
    ref = O.theta_grad(grid[:, sel].cpu().numpy(), As[:n_chk].cpu().numpy(), Bs,
                       g1[:n_chk][:, :, sel].cpu().numpy(), tess, 50, threads=8)
    got, _ = ops.backward_theta(grid[:, sel].contiguous(), As[:n_chk], B, g1[:n_chk][:, :, sel].contiguous(), tess, 50)
    assert_grad_parity(got.cpu().numpy(), ref, 1e-5)


def test_theta_sharding_is_exact():
    """Running two half-batches (what two ranks do) equals running the whole batch."""
    from libcpab_b200.distributed import shard
    torch.manual_seed(3)
    T = _T([3, 3])
    theta = T.sample_transformation(10)
    data = torch.rand(10, 1, 64, 64, device="cuda")
    whole = T.transform_data(data, theta, (64, 64))
    parts = [T.transform_data(shard(data, r, 3), shard(theta, r, 3), (64, 64)) for r in range(3)]
    assert bool((torch.cat(parts) == whole).all())
