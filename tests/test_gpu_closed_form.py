"""Opt-in closed-form (hit-time) 1-D integration -- the algorithm north_star describes, which the
reference does not contain.  Checked against the float64 numpy checker (oracle.closed_form_1d,
itself anchored to the reference's fixed-step semantics by convergence, see
tests/test_closed_form_oracle.py), against finite differences of that checker for the gradient,
and for flow properties at full size."""
import numpy as np
import pytest
import torch

from conftest import flow_gain, load_golden, rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t if dtype is None else t.to(dtype)).cuda()


@pytest.mark.parametrize("name,scale", [("cfg1_1d50", 1.0), ("d1_t100", 1.0), ("d1_t10_free", 0.5)])
def test_forward_matches_checker(name, scale):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"] * scale, nc, np.float64)
    grid = g["grid"].astype(np.float64)
    ref = O.closed_form_1d(grid, As, nc)
    got64 = ops.forward_closed_form(dev(grid), dev(As), nc).cpu().numpy()
    assert rel_err(got64, ref) < 1e-12
    got32 = ops.forward_closed_form(dev(grid, torch.float32), dev(As, torch.float32), nc).cpu().numpy()
    assert rel_err(got32, ref) < 1e-6 * flow_gain(As)
    # and it is what the reference's scheme converges to
    n = 5000
    approx = O.forward(grid, O.affine_to_trels(As, n), nc, n)
    assert np.abs(got64 - approx).max() < 2e-4


@pytest.mark.parametrize("name", ["cfg1_1d50", "d1_t10_free"])
def test_gradient_matches_differenced_checker(name):
    from libcpab_b200 import ops
    g = load_golden(name)
    nc = g["nc"].tolist()
    B = g["B"]
    theta = g["theta"][:2].astype(np.float64) * 0.7
    grid = g["grid"].astype(np.float64)[:, ::7]
    rng = np.random.default_rng(1)
    gout = rng.normal(size=(2, 1, grid.shape[1]))

    def loss(th):
        return float((O.closed_form_1d(grid, O.theta_to_affine(B, th, nc, np.float64), nc) * gout).sum())

    fd = np.zeros_like(theta)
    eps = 1e-6
    for t in range(theta.shape[0]):
        for k in range(theta.shape[1]):
            e = np.zeros_like(theta); e[t, k] = eps
            fd[t, k] = (loss(theta + e) - loss(theta - e)) / (2 * eps)
    As = O.theta_to_affine(B, theta, nc, np.float64)
    dth, dpts = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc, want_dpoints=True)
    assert rel_err(dth.cpu().numpy(), fd) < 1e-6
    # float32
    dth32, _ = ops.backward_theta_closed_form(dev(grid, torch.float32), dev(As, torch.float32),
                                              dev(B, torch.float32), dev(gout, torch.float32), nc)
    assert rel_err(dth32.cpu().numpy(), fd) < 2e-4
    # d/dpoints by differences of the checker
    fdp = (O.closed_form_1d(grid + eps, As, nc) - O.closed_form_1d(grid - eps, As, nc)) / (2 * eps) * gout
    assert rel_err(dpts.cpu().numpy(), fdp) < 1e-5


def test_gradient_at_the_identity_is_not_zero():
    """theta = 0 (the usual identity initialisation): every velocity vanishes, every point is a fixed
    point, and yet d x_f / d b = 1 and d x_f / d a = x.  The closed-form gradient there must equal the
    fixed-step adjoint's (which is exact for a field that is zero) -- ADVICE r1."""
    from libcpab_b200 import ops
    g = load_golden("cfg1_1d50")
    nc = g["nc"].tolist()
    B = g["B"]
    grid = g["grid"].astype(np.float64)
    As = np.zeros((3, nc[0], 1, 2))
    rng = np.random.default_rng(4)
    gout = rng.normal(size=(3, 1, grid.shape[1]))
    dth, dpts = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc, want_dpoints=True)
    ref, _ = ops.backward_theta(dev(grid), dev(As), dev(B), dev(gout), nc, 50)
    assert float(ref.abs().max()) > 1e-3
    assert rel_err(dth.cpu().numpy(), ref.cpu().numpy()) < 1e-9
    assert np.allclose(dpts.cpu().numpy(), gout)                          # d x_f / d x_0 = e^{a t} = 1
    # a training step from T.identity() moves
    from libcpab_b200 import Cpab
    T = Cpab(nc, backend="pytorch", device="gpu", basis=B)
    T.params.closed_form = True
    theta = T.identity(2).requires_grad_(True)
    out = T.transform_grid(T.uniform_meshgrid([64]), theta)
    (out * torch.linspace(-1, 1, 64, device="cuda")).sum().backward()
    assert float(theta.grad.abs().max()) > 1e-3


def test_api_switch_and_flow_properties_at_full_size():
    """BASELINE configs[4] shape: 8192 series x 1024 points, tess [100]."""
    from libcpab_b200 import Cpab
    torch.manual_seed(0)
    T = Cpab([100], backend="pytorch", device="gpu")
    T.params.closed_form = True
    theta = T.sample_transformation(8192).requires_grad_(True)
    grid = T.uniform_meshgrid([1024])
    out = T.transform_grid(grid, theta)
    assert tuple(out.shape) == (8192, 1, 1024) and bool(torch.isfinite(out).all())
    assert bool((out[:, 0, 1:] > out[:, 0, :-1]).all())                   # order preserving
    assert float((out[:, 0, 0]).abs().max()) < 1e-6 and float((out[:, 0, -1] - 1).abs().max()) < 1e-6
    back = T.transform_grid(out.detach(), -theta.detach())                 # inverse flow
    assert float((back - grid[None]).abs().max()) < 5e-5
    out.square().sum().backward()
    assert bool(torch.isfinite(theta.grad).all()) and float(theta.grad.abs().max()) > 0
    # close to the fixed-step result (which carries an O(1/N) crossing error)
    T.params.closed_form = False
    fixed = T.transform_grid(grid, theta.detach())
    assert float((fixed - out.detach()).abs().max()) < 5e-3
    # identity
    T.params.closed_form = True
    assert bool((T.transform_grid(grid, T.identity(3)) == grid[None]).all())


def test_closed_form_is_one_dimensional_only():
    from libcpab_b200 import Cpab, _lib, ops
    T = Cpab([3, 3], backend="pytorch", device="gpu")
    T.params.closed_form = True
    with pytest.raises(NotImplementedError):
        T.transform_grid(T.uniform_meshgrid([8, 8]), T.sample_transformation(2))
    with pytest.raises(_lib.CpabError, match="1-D only"):
        ops.forward_closed_form(torch.zeros(2, 8, device="cuda"), torch.zeros(1, 36, 2, 3, device="cuda"), [3, 3])
