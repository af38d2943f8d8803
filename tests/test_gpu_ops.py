"""GPU parity of every C-ABI op against the oracle and the committed reference fixtures.

Tolerances (north_star): cell indices bit-exact; points and gradients <= 1e-5 relative in
float32 (relative = max|x-ref| / max|ref|), <= 1e-10 in the float64 check mode.  The strict
(default) forward is held to the stronger bar it was designed for: bit-identical to the CPU
reference for identical Trels.
"""
import numpy as np
import pytest
import torch

from conftest import (assert_grad_parity, assert_grad_parity_fast_mode, bs_of, flow_gain, golden_cases,
                      load_golden, rel_err)
from oracle import oracle as O

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
F64_TOL = 1e-10


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------------------ cells
def test_findcellidx_golden_bit_exact():
    from libcpab_b200 import ops
    z = load_golden("cells")
    for key in [k for k in z.files if k.startswith("pts_")]:
        nc = [int(s) for s in key[4:].split("x")]
        got = ops.findcellidx(dev(z[key]), nc).cpu().numpy()
        assert np.array_equal(got, z["idx_" + key[4:]]), f"tessellation {nc}"


@pytest.mark.parametrize("nc", [[1], [50], [100], [3, 3], [10, 10], [7, 4], [4, 4, 4], [3, 2, 5]])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_findcellidx_random_and_adversarial(nc, dtype):
    from libcpab_b200 import ops
    rng = np.random.default_rng(len(nc) * 100 + nc[0])
    ndim, n = len(nc), 200_000
    lat = np.stack([rng.integers(0, 4 * nc[j] + 1, n) / (4.0 * nc[j]) for j in range(ndim)])
    latf = lat.astype(dtype)
    edge = rng.integers(0, 2, (ndim, n)) + rng.choice([0, 1e-9, -1e-9, 1e-7, -1e-7, 0.05, -0.05], (ndim, n))
    pts = np.concatenate([rng.uniform(-0.2, 1.2, (ndim, n)), rng.uniform(0, 1, (ndim, n)), lat,
                          np.nextafter(latf, dtype(2)), np.nextafter(latf, dtype(-2)), edge],
                         axis=1).astype(dtype)
    got = ops.findcellidx(dev(pts), nc).cpu().numpy()
    assert np.array_equal(got, O.findcellidx(pts, nc))


# ------------------------------------------------------------------------------------------ expm
def test_expm_matches_reference_pade13():
    import scipy.linalg
    from libcpab_b200 import ops
    z = load_golden("expm")
    for m in (2, 3, 4):
        A = z[f"A{m}"]
        truth = np.stack([scipy.linalg.expm(a) for a in A.astype(np.float64)])
        got = ops.expm(dev(A)).cpu().numpy()
        # evaluated in double and rounded once: half an ulp of the largest entry
        assert rel_err(got, truth) < 1.2e-7
        # the reference's float32 Pade (incl. its squaring error on the large-norm rows)
        assert rel_err(got, z[f"E{m}"]) < 2e-5
        got64 = ops.expm(dev(A.astype(np.float64))).cpu().numpy()
        assert rel_err(got64, truth) < 1e-12
        # the reference's float64 result carries float32-rounded Pade coefficients (expm.py:44):
        # 1.4e-7 off the truth on the large-norm rows, which the oracle reproduces exactly
        assert rel_err(O.expm_pade13(A.astype(np.float64)), z[f"E{m}_f64"]) < 1e-14
        assert rel_err(got64, z[f"E{m}_f64"]) < 3e-7


@pytest.mark.parametrize("name", golden_cases())
def test_theta_to_trels(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    Bt = dev(np.ascontiguousarray(g["B"].T), torch.float32)
    As, Tr = ops.theta_to_trels(dev(g["theta"]), Bt, nc, int(g["nstepsolver"]))
    assert rel_err(As.cpu().numpy(), g["As"]) < 1e-6
    assert np.abs(Tr.cpu().numpy() - g["Trels"]).max() < 3e-7       # entries are O(1)
    # float64 check mode against the numpy restatement
    Bt64 = dev(np.ascontiguousarray(g["B"].T))
    As64, Tr64 = ops.theta_to_trels(dev(g["theta"].astype(np.float64)), Bt64, nc, int(g["nstepsolver"]))
    Ao = O.theta_to_affine(g["B"], g["theta"], nc, np.float64)
    assert rel_err(As64.cpu().numpy(), Ao) < 1e-13
    assert np.abs(Tr64.cpu().numpy() - O.affine_to_trels(Ao, int(g["nstepsolver"]))).max() < 1e-12


# --------------------------------------------------------------------------------------- forward
@pytest.mark.parametrize("name", golden_cases())
def test_forward_strict_is_bit_identical_to_reference(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    got = ops.forward(dev(g["grid"]), dev(g["Trels"]), nc, int(g["nstepsolver"])).cpu().numpy()
    assert np.array_equal(got, g["grid_t"]), "max diff %g" % np.abs(got - g["grid_t"]).max()


@pytest.mark.parametrize("name", golden_cases())
def test_forward_fast_math_within_tolerance(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    got = ops.forward(dev(g["grid"]), dev(g["Trels"]), g["nc"].tolist(), int(g["nstepsolver"]),
                      fast_math=True).cpu().numpy()
    # FMA contraction changes the last bit of every step; the flow amplifies that (conftest.flow_gain)
    tol = F32_TOL * flow_gain(g["As"])
    if len(g["nc"]) == 3:
        # 3-D: points with a coordinate on/over the unit box hit the reference's discontinuous
        # branches (coord==1.0 quirk, push-inside); compare where the field is continuous
        keep = ((g["grid"] > 0.02) & (g["grid"] < 0.98)).all(axis=0)
        assert rel_err(got[:, :, keep], g["grid_t"][:, :, keep]) < tol
    else:
        assert rel_err(got, g["grid_t"]) < tol


@pytest.mark.parametrize("name", golden_cases())
def test_forward_fp64_check_mode(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"], nc, np.float64)
    Tr = O.affine_to_trels(As, int(g["nstepsolver"]))
    ref = O.forward(g["grid"].astype(np.float64), Tr, nc, int(g["nstepsolver"]))
    got = ops.forward(dev(g["grid"].astype(np.float64)), dev(Tr), nc, int(g["nstepsolver"])).cpu().numpy()
    assert rel_err(got, ref) < F64_TOL
    if len(nc) < 3 or nc[0] == nc[2]:
        # the reference's own numpy backend (scipy expm, float64); its 3-D cell search clamps z
        # with inc_z where the C++ uses inc_x, so it only agrees on cubic tessellations
        assert rel_err(got, g["grid_t_numpy64"]) < 1e-9


def test_forward_broadcast_grid_and_trace():
    """[n_theta,ndim,nP] grids (CpabSequential's 2nd+ warp) and per-step cell indices."""
    from libcpab_b200 import ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    grids = np.ascontiguousarray(g["grid_t"])                 # one grid per theta
    ref, cells = O.forward(grids, g["Trels"], nc, 50, trace=True)
    got = ops.forward(dev(grids), dev(g["Trels"]), nc, 50).cpu().numpy()
    assert np.array_equal(got, ref)
    # the first step's cell indices, bit-exact
    n_theta = grids.shape[0]
    for t in range(n_theta):
        idx = ops.findcellidx(dev(grids[t]), nc).cpu().numpy()
        assert np.array_equal(idx, cells[t, 0])


@pytest.mark.parametrize("nsteps", [1, 7, 50, 64])
def test_forward_other_step_counts(nsteps):
    from libcpab_b200 import ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"], nc)
    Tr = O.affine_to_trels(As, nsteps)
    got = ops.forward(dev(g["grid"]), dev(Tr), nc, nsteps).cpu().numpy()
    assert np.array_equal(got, O.forward(g["grid"], Tr, nc, nsteps))


def test_forward_ragged_sizes():
    """nP not a multiple of anything, n_theta = 1, nP = 1, empty batches."""
    from libcpab_b200 import ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    for nP in (1, 31, 257, 1023):
        pts = np.ascontiguousarray(g["grid"][:, :nP])
        got = ops.forward(dev(pts), dev(g["Trels"][:1]), nc, 50).cpu().numpy()
        assert np.array_equal(got, O.forward(pts, g["Trels"][:1], nc, 50))
    empty = ops.forward(dev(g["grid"][:, :0]), dev(g["Trels"]), nc, 50)
    assert tuple(empty.shape) == (g["Trels"].shape[0], 2, 0)
    none = ops.forward(dev(g["grid"]), dev(g["Trels"][:0]), nc, 50)
    assert none.shape[0] == 0


def test_forward_tuning_variants_are_equivalent():
    from libcpab_b200 import _lib, ops
    g = load_golden("d2_t10x10_vp")
    nc = g["nc"].tolist()
    base = ops.forward(dev(g["grid"]), dev(g["Trels"]), nc, 50).cpu().numpy()
    try:
        for key, val in (("fwd_ppt", 2), ("chunk_pts", 256), ("chunk_pts", 4096)):
            _lib.set_tuning(key, val)
            assert np.array_equal(ops.forward(dev(g["grid"]), dev(g["Trels"]), nc, 50).cpu().numpy(), base)
    finally:
        _lib.set_tuning("fwd_ppt", 1)
        _lib.set_tuning("chunk_pts", 1024)
        _lib.set_tuning("chunk_auto", 1)


# -------------------------------------------------------------------------------------- gradient
@pytest.mark.parametrize("name", [n for n in golden_cases() if "jac" in load_golden(n).files])
def test_jacobian_matches_reference_op(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    Bs = bs_of(g["B"], nc)
    got = ops.backward_jacobian(dev(g["grid"]), dev(g["As"]), dev(Bs), nc, 50).cpu().numpy()
    assert got.shape == g["jac"].shape
    assert np.array_equal(got, g["jac"]), "max diff %g" % np.abs(got - g["jac"]).max()


@pytest.mark.parametrize("name", golden_cases())
def test_backward_theta_matches_reference_gradient(name):
    """Default mode: every theta within 1e-5 of the reference's float32 gradient."""
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    redo = torch.zeros(g["As"].shape[0], dtype=torch.int32, device="cuda")
    dth, _ = ops.backward_theta(dev(g["grid"]), dev(g["As"]), dev(g["B"], torch.float32),
                                dev(g["gout"]), nc, 50, redo_count=redo)
    print("re-integrated trajectories: %d of %d" % (int(redo.sum()), g["As"].shape[0] * g["grid"].shape[-1]))
    assert_grad_parity(dth.cpu().numpy(), g["dtheta"], F32_TOL, what=name)


@pytest.mark.parametrize("name", golden_cases())
def test_backward_theta_fast_grad_mode(name):
    """Opt-in CPAB_FLAG_FAST_GRAD: no certificate, rare cell flips allowed (and bounded)."""
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    dth, _ = ops.backward_theta(dev(g["grid"]), dev(g["As"]), dev(g["B"], torch.float32),
                                dev(g["gout"]), nc, 50, fast_grad=True)
    assert_grad_parity_fast_mode(dth.cpu().numpy(), g["dtheta"], F32_TOL, what=name)


@pytest.mark.parametrize("name", golden_cases())
def test_rk2_cell_sequences_are_the_references(name):
    """The device trace of the adjoint's first pass against the oracle's RK2 trace
    (libcpab/core/cpab_ops.cpp:289-366): the reference-arithmetic mode reproduces EVERY cell of
    EVERY trajectory bit for bit; every trajectory the certificate passes has exactly those cells
    (so the default gradient never integrates along a neighbouring cell)."""
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    ref_cells, _ = O.rk2_trace(g["grid"], g["As"], nc, 50)
    strict, _ = ops.rk2_cell_trace(dev(g["grid"]), dev(g["As"]), nc, 50, mode=2)
    assert np.array_equal(strict.cpu().numpy(), ref_cells)
    cert, failed = ops.rk2_cell_trace(dev(g["grid"]), dev(g["As"]), nc, 50, mode=1)
    cert, failed = cert.cpu().numpy(), failed.cpu().numpy().astype(bool)
    differs = (cert != ref_cells).any(axis=1)                       # [n_theta, nP]
    print("%s: certificate failed on %.2f %% of the trajectories; %d uncertified cell sequences differ, "
          "%d certified ones" % (name, 100.0 * failed.mean(), int((differs & failed).sum()), int((differs & ~failed).sum())))
    assert not (differs & ~failed).any()


@pytest.mark.parametrize("name", ["cfg1_1d50", "d2_t3x3", "d3_t2x2x2", "d2_t2x3_free_vp"])
def test_backward_theta_fp64_check_mode(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    n_theta = min(2, g["theta"].shape[0])
    As = O.theta_to_affine(g["B"], g["theta"][:n_theta], nc, np.float64)
    grid = g["grid"].astype(np.float64)
    gout = g["gout"][:n_theta].astype(np.float64)
    ref = O.theta_grad(grid, As, bs_of(g["B"], nc, np.float64), gout, nc, 50, threads=8)
    dth, _ = ops.backward_theta(dev(grid), dev(As), dev(g["B"]), dev(gout), nc, 50)
    # the adjoint restates the same RK2 recursion; it differs from the per-k form only by
    # floating-point association (and by `h` products the reference rounds through float)
    assert rel_err(dth.cpu().numpy(), ref) < F64_TOL


def test_backward_dpoints_against_differenced_flow():
    """dL/dpoints (extension; the reference returns None) against central differences of the
    oracle's float64 RK2 flow, on points whose trajectories stay clear of cell boundaries."""
    from libcpab_b200 import ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"][:2], nc, np.float64) * 0.5
    rng = np.random.default_rng(0)
    pts = rng.uniform(0.02, 0.98, (2, 400))
    gout = rng.normal(size=(2, 2, 400))
    _, dp = ops.backward_theta(dev(pts), dev(As), dev(g["B"]), dev(gout), nc, 50, want_dpoints=True)
    dp = dp.cpu().numpy()
    eps = 1e-6
    fd = np.zeros_like(dp)
    for j in range(2):
        e = np.zeros((2, 1)); e[j] = eps
        fp, fm = O.rk2_flow(pts + e, As, nc), O.rk2_flow(pts - e, As, nc)
        fd[:, j] = ((fp - fm) / (2 * eps) * gout).sum(axis=1)
    # a trajectory that crosses a cell boundary has a kink the adjoint (like the reference's own
    # gradient) ignores; compare where +-eps probes and the centre visit the same cells
    _, c0 = O.forward(pts, O.affine_to_trels(As), nc, 50, trace=True)
    same = (c0 == c0[:, :1]).all(axis=1)                    # never left the starting cell
    assert same.mean() > 0.2
    err = np.abs(dp - fd)[np.broadcast_to(same[:, None, :], dp.shape)]
    assert err.max() < 1e-6 * max(1.0, np.abs(fd).max())


def test_backward_other_step_counts_and_tuning():
    from libcpab_b200 import _lib, ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    B32 = dev(g["B"], torch.float32)
    for nsteps in (1, 7, 23, 50, 101):
        ref = O.theta_grad(g["grid"], g["As"], bs_of(g["B"], nc), g["gout"], nc, nsteps, threads=8)
        for seg, block in ((10, 128), (5, 64), (5, 256), (3, 128)):
            try:
                _lib.set_tuning("bwd_seg", seg)
                _lib.set_tuning("bwd_block", block)
                dth, _ = ops.backward_theta(dev(g["grid"]), dev(g["As"]), B32, dev(g["gout"]), nc, nsteps)
                dth_fast, _ = ops.backward_theta(dev(g["grid"]), dev(g["As"]), B32, dev(g["gout"]), nc, nsteps, fast_grad=True)
            finally:
                _lib.set_tuning("bwd_seg", 0)
                _lib.set_tuning("bwd_block", 128)
            assert_grad_parity(dth.cpu().numpy(), ref, F32_TOL, what="nsteps=%d seg=%d block=%d" % (nsteps, seg, block))
            assert_grad_parity_fast_mode(dth_fast.cpu().numpy(), ref, F32_TOL)


@pytest.mark.parametrize("nc,n_theta,nP", [([7], 1, 1), ([7], 3, 31), ([50], 700, 257), ([3, 3], 1, 255),
                                           ([3, 3], 5, 1000), ([4, 3], 2, 5000), ([2, 2, 2], 3, 513)])
def test_backward_ragged_sizes_through_the_work_units(nc, n_theta, nP):
    """Work units of the adjoint kernel (bulk units, small tail units, partial last units, more
    or fewer units than resident CTAs): float64 check mode against the oracle, so that no cell
    flip can hide an indexing error."""
    from libcpab_b200 import _lib, ops
    rng = np.random.default_rng(nP + n_theta)
    ndim = len(nc)
    nC = int({1: 1, 2: 4, 3: 5}[ndim] * np.prod(nc))
    d = 4
    B = rng.normal(size=(nC * ndim * (ndim + 1), d)) * 0.4
    theta = rng.normal(size=(n_theta, d))
    As = O.theta_to_affine(B, theta, nc, dtype=np.float64)
    pts = rng.uniform(0.0, 1.0, (ndim, nP))
    gout = rng.normal(size=(n_theta, ndim, nP))
    ref = O.theta_grad(pts, As, bs_of(B, nc, np.float64), gout, nc, 50, threads=8)
    for chunk in (None, 256, 4096):
        try:
            if chunk:
                _lib.set_tuning("chunk_pts", chunk)
            dth, dpts = ops.backward_theta(dev(pts), dev(As), dev(B), dev(gout), nc, 50, want_dpoints=True)
        finally:
            _lib.set_tuning("chunk_pts", 1024)
            _lib.set_tuning("chunk_auto", 1)
        assert rel_err(dth.cpu().numpy(), ref) < 1e-9, chunk
        assert dpts.shape == (n_theta, ndim, nP) and bool(torch.isfinite(dpts).all())


# --------------------------------------------------------------------------------- interpolation
@pytest.mark.parametrize("name", [n for n in golden_cases() if "data" in load_golden(n).files])
def test_interpolate_forward_bit_identical(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    out = ops.interpolate_forward(dev(g["data"]), dev(g["grid_t"]), g["grid_n"].tolist()).cpu().numpy()
    assert out.shape == g["interp_out"].shape
    assert np.array_equal(out, g["interp_out"]), "max diff %g" % np.abs(out - g["interp_out"]).max()


@pytest.mark.parametrize("name", [n for n in golden_cases() if "data" in load_golden(n).files])
def test_interpolate_backward(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    dgrid, ddata = ops.interpolate_backward(dev(g["data"]), dev(g["grid_t"]), dev(g["data_gout"]),
                                            want_dgrid=True, want_ddata=True)
    assert rel_err(dgrid.cpu().numpy(), g["interp_dgrid"]) < F32_TOL
    dg_o, dd_o = O.interpolate_vjp(g["data"], g["grid_t"], g["grid_n"].tolist(), g["data_gout"])
    assert rel_err(ddata.cpu().numpy(), dd_o) < F32_TOL
    assert rel_err(dgrid.cpu().numpy(), dg_o) < F32_TOL


@pytest.mark.parametrize("shape,outsize", [((3, 2, 37), (53,)), ((2, 3, 19, 45), (33, 70)),
                                           ((2, 2, 9, 11, 13), (17, 5, 40))])
def test_interpolate_ragged_shapes_and_outside_points(shape, outsize):
    from libcpab_b200 import ops
    rng = np.random.default_rng(sum(shape))
    ndim = len(shape) - 2
    data = rng.uniform(size=shape).astype(np.float32)
    grid = rng.uniform(-0.3, 1.3, (shape[0], ndim, int(np.prod(outsize)))).astype(np.float32)
    out = ops.interpolate_forward(dev(data), dev(grid), outsize).cpu().numpy()
    assert np.array_equal(out, O.interpolate(data, grid, outsize))
    gout = rng.normal(size=out.shape).astype(np.float32)
    dgrid, ddata = ops.interpolate_backward(dev(data), dev(grid), dev(gout), True, True)
    dg_o, dd_o = O.interpolate_vjp(data, grid, outsize, gout)
    assert rel_err(dgrid.cpu().numpy(), dg_o) < F32_TOL
    assert rel_err(ddata.cpu().numpy(), dd_o) < F32_TOL
    # float64 check mode
    out64 = ops.interpolate_forward(dev(data.astype(np.float64)), dev(grid.astype(np.float64)), outsize)
    assert rel_err(out64.cpu().numpy(), O.interpolate(data.astype(np.float64), grid.astype(np.float64), outsize)) < 1e-14


@pytest.mark.parametrize("shape,outsize,cap", [((5, 1, 40, 33), (70, 100), 3), ((3, 2, 21, 30), (65, 129), 7),
                                               ((2, 1, 9, 11, 13), (40, 6, 70), 2), ((1, 2, 12, 7, 9), (33, 3, 33), 0)])
def test_interpolate_persistent_kernels_walk_many_tiles(shape, outsize, cap):
    """The float32 interpolate kernels are persistent (a CTA walks over 32x32 tiles with the grid
    tiles of the next two iterations in flight).  With the grid capped to a few CTAs every CTA
    goes around its 3-stage ring several times, over interior and ragged edge tiles alike."""
    from libcpab_b200 import _lib, ops
    rng = np.random.default_rng(sum(shape) + cap)
    ndim = len(shape) - 2
    data = rng.uniform(size=shape).astype(np.float32)
    grid = rng.uniform(-0.2, 1.2, (shape[0], ndim, int(np.prod(outsize)))).astype(np.float32)
    gout = rng.normal(size=(shape[0], shape[1], *outsize)).astype(np.float32)
    ref = O.interpolate(data, grid, outsize)
    dg_o, dd_o = O.interpolate_vjp(data, grid, outsize, gout)
    try:
        _lib.set_tuning("interp_max_ctas", cap)
        for var in (2, 5, 6, 7, 8, 9, 10, 11):
            _lib.set_tuning("interp_variant", var)
            out = ops.interpolate_forward(dev(data), dev(grid), outsize).cpu().numpy()
            assert np.array_equal(out, ref), var
            dgrid, ddata = ops.interpolate_backward(dev(data), dev(grid), dev(gout), True, True)
            assert rel_err(dgrid.cpu().numpy(), dg_o) < F32_TOL, var
            assert rel_err(ddata.cpu().numpy(), dd_o) < F32_TOL, var
            dgrid2, none = ops.interpolate_backward(dev(data), dev(grid), dev(gout), True, False)
            assert none is None and torch.equal(dgrid2, dgrid), var
    finally:
        _lib.set_tuning("interp_max_ctas", 0)
        _lib.set_tuning("interp_variant", 9)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_interpolate_batch_beyond_the_grid_z_limit(dtype):
    """More than 65535 samples through the one-tile-per-CTA kernels (float64, and float32 with an
    output extent the 16-byte path does not take): the batch is slabbed, as the reference has no limit."""
    from libcpab_b200 import ops
    rng = np.random.default_rng(9)
    N, outsize = 66000, (5, 3)
    data = rng.uniform(size=(N, 1, 4, 3)).astype(dtype)
    grid = rng.uniform(-0.1, 1.1, (N, 2, 15)).astype(dtype)
    out = ops.interpolate_forward(dev(data), dev(grid), outsize).cpu().numpy()
    sel = np.r_[0:50, 65500:65600, N - 50:N]
    assert np.array_equal(out[sel], O.interpolate(data[sel], grid[sel], outsize))
    gout = rng.normal(size=out.shape).astype(dtype)
    dgrid, ddata = ops.interpolate_backward(dev(data), dev(grid), dev(gout), True, True)
    dg_o, dd_o = O.interpolate_vjp(data[sel], grid[sel], outsize, gout[sel])
    tol = F32_TOL if dtype == np.float32 else 1e-12
    assert rel_err(dgrid.cpu().numpy()[sel], dg_o) < tol
    assert rel_err(ddata.cpu().numpy()[sel], dd_o) < tol


def test_interpolate_taps_extreme_coordinates():
    """The conversion-free tap arithmetic (cpab_sample.cuh) against the oracle on coordinates far
    outside the image, exactly on texels, just below them, negative zero, and huge (below 2^63 after
    scaling: beyond, the reference's float -> int64 conversion is itself undefined)."""
    from libcpab_b200 import ops
    W, H = 17, 9
    rng = np.random.default_rng(3)
    data = rng.uniform(size=(1, 1, W, H)).astype(np.float32)
    xs = np.array([0.0, -0.0, 1.0, np.nextafter(np.float32(1), np.float32(0)), np.nextafter(np.float32(1), np.float32(2)),
                   -1e-8, 1e-8, 0.5, 1.0 / 16, 3.0 / 16 - 1e-7, -0.07, -1.0, -3.5, 2.0, 7.25, 1e6, -1e6, 3e9, -3e9],
                  dtype=np.float32)
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    n = xs.size
    grid = np.stack([gx.ravel(), gy.ravel()])[None].astype(np.float32)
    out = ops.interpolate_forward(dev(data), dev(grid), [n, n]).cpu().numpy()
    ref = O.interpolate(data, grid, [n, n])
    assert np.array_equal(out, ref, equal_nan=True)


# ------------------------------------------------------------------------------------- error path
def test_errors_are_raised_not_swallowed():
    from libcpab_b200 import _lib, ops
    g = load_golden("d2_t3x3")
    with pytest.raises(RuntimeError):
        ops.forward(torch.from_numpy(g["grid"]), dev(g["Trels"]), [3, 3], 50)      # CPU tensor
    with pytest.raises(ValueError):
        ops.forward(dev(g["grid"]), dev(g["Trels"]), [4, 4], 50)                   # wrong tessellation
    with pytest.raises(_lib.CpabError):
        ops.forward(dev(g["grid"]), dev(g["Trels"]), [3, 3], 0)                    # nstepsolver = 0
    with pytest.raises(_lib.CpabError):
        _lib.set_tuning("no_such_knob", 1)


# ------------------------------------------------------------------- large tessellations / step counts
def test_tessellation_too_large_for_shared_memory():
    """[128,130] -> 66 560 simplices: matrices are read through L1 instead of staged, the cell
    trace is 32-bit.  Random (not constraint-satisfying) fields: the kernels do not care."""
    from libcpab_b200 import ops
    rng = np.random.default_rng(5)
    nc = [128, 130]
    nC, d, n_theta, nP = 4 * 128 * 130, 3, 2, 3000
    B = rng.normal(size=(nC * 6, d)).astype(np.float32) * 0.3
    theta = rng.normal(size=(n_theta, d)).astype(np.float32)
    As = O.theta_to_affine(B, theta, nc)
    Tr = O.affine_to_trels(As)
    pts = rng.uniform(-0.05, 1.05, (2, nP)).astype(np.float32)
    got = ops.forward(dev(pts), dev(Tr), nc, 50).cpu().numpy()
    assert np.array_equal(got, O.forward(pts, Tr, nc, 50))
    gout = rng.normal(size=(n_theta, 2, nP)).astype(np.float32)
    ref = O.theta_grad(pts, As, bs_of(B, nc), gout, nc, 50, threads=8)
    dth, _ = ops.backward_theta(dev(pts), dev(As), dev(B), dev(gout), nc, 50)
    assert_grad_parity(dth.cpu().numpy(), ref, F32_TOL)
    As_g, Tr_g = ops.theta_to_trels(dev(theta), dev(np.ascontiguousarray(B.T)), nc, 50)
    assert rel_err(As_g.cpu().numpy(), As) < 1e-6


def test_many_solver_steps_and_the_limit():
    from libcpab_b200 import _lib, ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    pts = np.ascontiguousarray(g["grid"][:, ::5])
    gout = np.ascontiguousarray(g["gout"][:, :, ::5])
    B32 = dev(g["B"], torch.float32)
    for nsteps in (400, 1000):
        ref = O.theta_grad(pts, g["As"], bs_of(g["B"], nc), gout, nc, nsteps, threads=8)
        dth, _ = ops.backward_theta(dev(pts), dev(g["As"]), B32, dev(gout), nc, nsteps)
        assert_grad_parity(dth.cpu().numpy(), ref, F32_TOL)
    with pytest.raises(_lib.CpabError, match="checkpoint memory"):
        ops.backward_theta(dev(pts), dev(g["As"]), B32, dev(gout), nc, 20000)
    # the forward has no such limit
    Tr = O.affine_to_trels(g["As"], 20000)
    out = ops.forward(dev(pts), dev(Tr), nc, 20000).cpu().numpy()
    assert np.array_equal(out, O.forward(pts, Tr, nc, 20000))


def test_non_finite_inputs_do_not_crash():
    from libcpab_b200 import ops
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    pts = g["grid"][:, :64].copy()
    pts[0, 3], pts[1, 7], pts[0, 11], pts[1, 12] = np.nan, np.inf, -np.inf, 1e30
    out = ops.forward(dev(pts), dev(g["Trels"]), nc, 50).cpu().numpy()
    keep = np.isfinite(pts).all(axis=0) & (np.abs(pts) < 10).all(axis=0)
    assert np.array_equal(out[:, :, keep], O.forward(pts, g["Trels"], nc, 50)[:, :, keep])
    idx = ops.findcellidx(dev(pts), nc).cpu().numpy()
    assert idx.min() >= 0 and idx.max() < 36
