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
