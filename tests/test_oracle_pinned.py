"""Pins the CPU oracle (oracle/) to the reference: every function is checked against fixtures the
UNMODIFIED reference produced (tests/golden/make_golden.py) and, when present, against the
reference's own object code (oracle/_ref/libcpab_ref.so).  Runs on the CPU."""
import numpy as np
import pytest

from conftest import bs_of, golden_cases, load_golden, rel_err
from oracle import oracle as O

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference)")


def test_cell_indices_match_reference_extension():
    z = load_golden("cells")
    keys = [k for k in z.files if k.startswith("pts_")]
    assert len(keys) == 13
    for key in keys:
        nc = [int(s) for s in key[4:].split("x")]
        assert np.array_equal(O.findcellidx(z[key], nc), z["idx_" + key[4:]]), nc


@needs_ref
def test_cell_indices_match_reference_object_code():
    rng = np.random.default_rng(1)
    for nc in ([50], [3, 3], [10, 10], [5, 2], [4, 4, 4], [3, 2, 5]):
        n, ndim = 50_000, len(nc)
        lat = np.stack([rng.integers(0, 4 * nc[j] + 1, n) / (4.0 * nc[j]) for j in range(ndim)])
        pts = np.concatenate([rng.uniform(-0.2, 1.2, (ndim, n)), lat,
                              np.nextafter(lat.astype(np.float32), np.float32(2))], axis=1).astype(np.float32)
        assert np.array_equal(O.findcellidx(pts, nc), O.ref_findcellidx(pts, nc)), nc


def test_expm_restatement_matches_reference_torch_expm():
    z = load_golden("expm")
    for m in (2, 3, 4):
        assert rel_err(O.expm_pade13(z[f"A{m}"]), z[f"E{m}"]) < 2e-6          # float32, own rounding
        assert rel_err(O.expm_pade13(z[f"A{m}"].astype(np.float64)), z[f"E{m}_f64"]) < 1e-14


@pytest.mark.parametrize("name", golden_cases())
def test_forward_jacobian_gradient_match_reference(name):
    g = load_golden(name)
    nc = g["nc"].tolist()
    n = int(g["nstepsolver"])
    # host pieces
    As = O.theta_to_affine(g["B"], g["theta"], nc)
    assert rel_err(As, g["As"]) < 1e-6
    assert np.abs(O.affine_to_trels(As, n) - g["Trels"]).max() < 2.5e-7      # <= 2 ulp at 1.0
    # native core, identical inputs -> identical bits
    assert np.array_equal(O.forward(g["grid"], g["Trels"], nc, n), g["grid_t"])
    Bs = bs_of(g["B"], nc)
    if "jac" in g.files:
        assert np.array_equal(O.jacobian(g["grid"], g["As"], Bs, nc, n, threads=4), g["jac"])
    dth = O.theta_grad(g["grid"], g["As"], Bs, g["gout"], nc, n, threads=8)
    assert rel_err(dth, g["dtheta"]) < 1e-6        # torch sums the products in float32


@needs_ref
@pytest.mark.parametrize("name", ["d1_t100", "d2_t3x3", "d3_t2x2x2", "d2_t2x3_free_vp"])
def test_restatement_equals_reference_object_code(name):
    g = load_golden(name)
    nc = g["nc"].tolist()
    assert np.array_equal(O.forward(g["grid"], g["Trels"], nc, 50), O.ref_forward(g["grid"], g["Trels"], nc, 50, threads=2))
    Bs = bs_of(g["B"], nc)
    assert np.array_equal(O.jacobian(g["grid"], g["As"], Bs, nc, 50, threads=4),
                          O.ref_jacobian(g["grid"], g["As"], Bs, nc, 50, threads=3))


@pytest.mark.parametrize("name", ["cfg1_1d50", "d1_t10_free", "d2_t3x3", "d2_t2x3_free_vp", "d3_t2x2x2"])
def test_fp64_variant_against_reference_numpy_backend(name):
    """The all-double oracle vs the reference's float64 numpy backend (scipy expm).  Valid where
    the two cell searches share semantics: 1-D, 2-D, and cubic 3-D tessellations away from the
    coord==1.0 faces (the numpy search clamps with n*inc-1e-8 in double)."""
    g = load_golden(name)
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"], nc, np.float64)
    out = O.forward(g["grid"].astype(np.float64), O.affine_to_trels(As), nc, 50)
    keep = np.ones(g["grid"].shape[1], bool)
    if len(nc) == 3:
        keep = ((g["grid"] > 0.01) & (g["grid"] < 0.99)).all(axis=0)
    assert rel_err(out[:, :, keep], g["grid_t_numpy64"][:, :, keep]) < 1e-10


@pytest.mark.parametrize("name", [n for n in golden_cases() if "data" in load_golden(n).files])
def test_interpolation_and_its_vjp_match_reference(name):
    g = load_golden(name)
    outsize = g["grid_n"].tolist()
    assert np.array_equal(O.interpolate(g["data"], g["grid_t"], outsize), g["interp_out"])
    dgrid, _ = O.interpolate_vjp(g["data"], g["grid_t"], outsize, g["data_gout"])
    assert rel_err(dgrid, g["interp_dgrid"]) < 2e-6
    # transform_data = meshgrid -> forward -> interpolate; its data-gradient is the scatter VJP
    grid0 = O.uniform_meshgrid(outsize)
    gt = O.forward(grid0, g["Trels"], g["nc"].tolist(), 50)
    assert np.array_equal(O.interpolate(g["data"], gt, outsize), g["data_t"])
    _, ddata = O.interpolate_vjp(g["data"], gt, outsize, g["data_gout"])
    # (outside the domain both taps clamp to the same texel with weights like 4.2 and -3.2; the
    #  reference adds those two float32 contributions, the oracle sums in double)
    assert rel_err(ddata, g["data_ddata"]) < 5e-5
    # and the theta-gradient of the composition
    dg, _ = O.interpolate_vjp(g["data"], gt, outsize, g["data_gout"])
    dth = O.theta_grad(grid0, g["As"], bs_of(g["B"], g["nc"].tolist()), dg, g["nc"].tolist(), 50, threads=8)
    assert rel_err(dth, g["data_dtheta"]) < 5e-6


def test_meshgrid_ordering():
    import torch
    for n in ([7], [5, 4], [3, 4, 5]):
        lin = [torch.linspace(0, 1, k) for k in n]
        mesh = torch.meshgrid(lin[::-1], indexing="ij")
        ref = torch.cat([m.reshape(1, -1) for m in mesh[::-1]], 0).numpy()
        assert np.array_equal(O.uniform_meshgrid(n), ref)


def test_sequential_fixture_is_reproduced():
    g = load_golden("seq_1d20x3")
    grid = O.uniform_meshgrid([128])
    for i in range(3):
        As = O.theta_to_affine(g["B"], g[f"theta{i}"], [20])
        grid = O.forward(grid, O.affine_to_trels(As), [20], 50)
        assert rel_err(grid, g["grids"][i]) < 2e-4      # own float32 expm rounding, chained flows
    assert not np.any(g["dtheta0"]) and not np.any(g["dtheta1"]) and np.any(g["dtheta2"])


def test_reference_gradient_moves_more_than_1e5_between_float32_and_float64():
    """Evidence for conftest.assert_grad_parity: the reference's arithmetic (oracle, pinned above
    bit for bit to the reference's float32 Jacobian) evaluated in float32 and in float64 on
    BASELINE configs[0] gives theta-gradients that differ by far more than 1e-5 of the largest
    entry, and the difference is concentrated in a few thetas -- cell assignments of near-face
    iterates flip with the rounding, everything else agrees to rounding."""
    from conftest import theta_errs
    g = load_golden("cfg1_1d50")
    nc = g["nc"].tolist()
    g32 = O.theta_grad(g["grid"], g["As"], bs_of(g["B"], nc), g["gout"], nc, 50, threads=8)
    assert rel_err(g32, g["dtheta"]) < 2e-7                      # oracle == reference (float32)
    f8 = np.float64
    g64 = O.theta_grad(g["grid"].astype(f8), g["As"].astype(f8), bs_of(g["B"], nc, f8), g["gout"].astype(f8),
                       nc, 50, threads=8)
    errs = theta_errs(g32, g64)
    assert errs.max() > 2e-5, errs.max()
    assert np.median(errs) < 5e-6, np.median(errs)


# ------------------------------------------------------------------ prior covariance / vector field
def _prior_cov(B, centers, ppc, length_scale, output_variance):
    """libcpab/cpab.py:211-237 restated: squared-exponential kernel on squared centre distances on
    the per-parameter diagonal of every (cell, cell) block, 100 * max distance elsewhere."""
    c = np.asarray(centers, dtype=np.float64)
    n2 = (c * c).sum(1)[:, None]
    dist = n2 - 2 * c @ c.T + n2.T
    eye = np.eye(ppc)
    cov_init = np.kron(dist, eye) + np.kron(np.ones_like(dist), 100 * dist.max() * (1 - eye))
    cov_avees = output_variance ** 2 * np.exp(-(cov_init / (2 * length_scale ** 2)))
    return B.T @ cov_avees @ B


@pytest.mark.parametrize("name", __import__("conftest").aux_cases())
def test_prior_covariance_and_vectorfield_restatements_match_reference(name):
    """The reference's smooth prior (its block loop, run by make_golden_aux.py) equals the Kronecker
    restatement the product uses; its calc_vectorfield equals A[cell(p)] [p;1] with the oracle's cells;
    the product's cell centres are the reference's."""
    from libcpab_b200.tessellation import cell_vertices
    g = load_golden(name)
    nc = g["nc"].tolist()
    ndim = len(nc)
    ppc = ndim * (ndim + 1)
    cov = _prior_cov(g["B"], g["centers"], ppc, float(g["length_scale"]), float(g["output_variance"]))
    assert rel_err(cov, g["cov_theta"]) < 1e-12
    centers = np.mean(cell_vertices(nc)[:, :, :ndim], axis=1)        # Tessellation.get_cell_centers
    assert np.allclose(centers, g["centers"], rtol=0, atol=1e-15)
    As = (g["B"].astype(np.float32) @ g["theta"][0].astype(np.float32)).reshape(-1, ndim, ndim + 1)
    cells = O.findcellidx(g["grid"], nc)
    homog = np.concatenate([g["grid"], np.ones((1, g["grid"].shape[1]), np.float32)])
    v = np.einsum("pij,jp->ip", As[cells], homog)
    # zero-boundary fields cancel (|a x|, |b| >> |v| = |a x + b|): float32 evaluation order shows at 1e-5
    assert rel_err(v, g["vectorfield"]) < 2e-5
