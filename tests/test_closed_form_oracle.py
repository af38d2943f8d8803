"""The closed-form 1-D checker (oracle.closed_form_1d) against the reference's semantics: the
reference has no closed-form integrator, so the anchor is convergence -- its fixed-step scheme
(restated by oracle.forward, float64) tends to the exact flow as nstepsolver grows."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O


@pytest.mark.parametrize("name,scale", [("cfg1_1d50", 1.0), ("d1_t100", 1.0), ("d1_t10_free", 0.5)])
def test_fixed_step_scheme_converges_to_the_closed_form(name, scale):
    g = load_golden(name)
    nc = g["nc"].tolist()
    theta = g["theta"][:4] * scale
    As = O.theta_to_affine(g["B"], theta, nc, np.float64)
    grid = g["grid"].astype(np.float64)
    exact = O.closed_form_1d(grid, As, nc)
    errs = []
    for n in (50, 500, 5000):
        approx = O.forward(grid, O.affine_to_trels(As, n), nc, n)
        errs.append(np.abs(approx - exact).max())
    assert errs[1] < errs[0] and errs[2] < errs[1]            # monotone convergence
    assert errs[2] < 20 * errs[0] / 100                        # ~ first order in 1/N
    assert errs[2] < 2e-4


def test_closed_form_is_a_flow():
    """phi(theta) followed by phi(-theta) is the identity; phi(0) is the identity."""
    g = load_golden("d1_t100")
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"][:3], nc, np.float64)
    grid = g["grid"].astype(np.float64)
    fwd = O.closed_form_1d(grid, As, nc)
    back = O.closed_form_1d(fwd, -As, nc)
    assert np.abs(back - grid[None]).max() < 1e-12
    assert np.abs(O.closed_form_1d(grid, 0 * As, nc) - grid[None]).max() == 0.0
    assert np.all(np.diff(fwd, axis=-1) > 0)                   # diffeomorphism: order preserving


# ---- 2-D / 3-D: oracle.closed_form_nd ---------------------------------------------------------------
ND_CASES = [("d2_t3x3", 60), ("d2_t10x10_vp", 60), ("d3_t2x2x2", 60), ("d2_t2x3_free_vp", 40), ("d3_t2x2x2_free", 40)]


@pytest.mark.parametrize("name,npts", ND_CASES)
def test_nd_rk2_flow_converges_to_the_hit_time_flow(name, npts):
    """Anchor of the n-D checker: the float64 RK2 flow of the same field (the scheme of
    cpab_ops.cpp:289-366, restated in double by the pinned C oracle) tends to it at second order.
    Trajectories that leave the unit box (tessellations without zero boundary) are left out: the
    reference's treatment of outside points is its own (cpab_ops.cpp:47-92, 119-136)."""
    g = load_golden(name)
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"][-2:], nc, np.float64)
    grid = g["grid"].astype(np.float64)
    grid = grid[:, ::max(1, grid.shape[1] // npts)]
    st = {}
    exact = O.closed_form_nd(grid, As, nc, st)
    inside = ~st["outside"]
    assert inside.sum() >= 20
    errs = []
    for n in (200, 2000, 20000):
        approx = O.rk2_flow(grid, As, nc, n)
        errs.append(np.abs(approx - exact).max(axis=1)[inside].max())
    print(name, "RK2 error at 200 / 2000 / 20000 steps:", errs, "sub-steps per trajectory: %.1f" % (st["segments"] / inside.size))
    assert errs[1] < errs[0] / 30 and errs[2] < errs[1] / 30          # ~ second order in 1/N
    assert errs[2] < 1e-9


@pytest.mark.parametrize("name", ["d2_t3x3", "d3_t2x2x2"])
def test_nd_hit_time_flow_is_a_flow(name):
    g = load_golden(name)
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"][:2], nc, np.float64)
    grid = g["grid"].astype(np.float64)
    grid = grid[:, ::max(1, grid.shape[1] // 40)]
    fwd = O.closed_form_nd(grid, As, nc)
    back = O.closed_form_nd(fwd, -As, nc)
    assert np.abs(back - grid[None]).max() < 1e-11
    assert np.abs(O.closed_form_nd(grid, 0 * As, nc) - grid[None]).max() == 0.0
    assert np.abs(fwd - grid[None]).max() > 1e-2


def test_nd_face_tables_hold_for_every_fixture_tessellation():
    import oracle.oracle as OO
    for name in ("d2_t3x3", "d2_t10x10_vp", "d2_t2x3_free_vp", "d3_t2x2x2", "d3_t3x2x2_vp"):
        OO._check_geometry([int(v) for v in load_golden(name)["nc"]])
