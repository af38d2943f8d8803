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


# ---- 2-D / 3-D: the hit-time walk on Taylor polynomials (cpab_closednd.cu) ----------------------------
ND = [("d2_t3x3", 150), ("d2_t10x10_vp", 150), ("d3_t2x2x2", 120), ("d3_t3x2x2_vp", 60), ("d2_t2x3_free_vp", 80),
      ("d3_t2x2x2_free", 80)]


def _nd_case(name, npts, n_theta=3):
    g = load_golden(name)
    nc = g["nc"].tolist()
    theta = g["theta"][-n_theta:].astype(np.float64)
    As = O.theta_to_affine(g["B"], theta, nc, np.float64)
    grid = g["grid"].astype(np.float64)
    grid = grid[:, ::max(1, grid.shape[1] // npts)]
    return g, nc, theta, As, grid


@pytest.mark.parametrize("name,npts", ND)
def test_nd_forward_matches_checker(name, npts):
    """Kernel (Taylor polynomials, Newton on the face polynomials) against the checker (scipy expm,
    brentq): float64 to 1e-10, float32 to 1e-5 relative -- every trajectory, including the ones of the
    free-boundary tessellations that leave the unit box (both continue the boundary cubes' planes)."""
    from libcpab_b200 import ops
    g, nc, theta, As, grid = _nd_case(name, npts)
    st = {}
    ref = O.closed_form_nd(grid, As, nc, st)
    got64 = ops.forward_closed_form(dev(grid), dev(As), nc).cpu().numpy()
    e64 = np.abs(got64 - ref).max()
    got32 = ops.forward_closed_form(dev(grid, torch.float32), dev(As, torch.float32), nc).cpu().numpy()
    e32 = np.abs(got32 - ref).max()
    print("%s: float64 max |err| %.2e, float32 %.2e (bound %.1e), %d of %d trajectories leave the box, %.1f sub-steps each"
          % (name, e64, e32, 1e-6 * flow_gain(As), st["outside"].sum(), st["outside"].size, st["segments"] / st["outside"].size))
    assert e64 < 1e-10
    assert e32 < 1e-6 * flow_gain(As) + 2e-6


@pytest.mark.parametrize("name,npts", [("d2_t3x3", 64), ("d3_t2x2x2", 64), ("d2_t10x10_vp", 64)])
def test_nd_gradient_matches_differenced_checker(name, npts):
    """d/dtheta and d/dpoints: float64 kernel against central differences of the CHECKER along random
    directions, and against central differences of the kernel's own float64 forward for every theta
    component; float32 kernel against the float64 one."""
    from libcpab_b200 import ops
    g, nc, theta, As, grid = _nd_case(name, npts, n_theta=2)
    B = g["B"]
    rng = np.random.default_rng(3)
    gout = rng.normal(size=(theta.shape[0], len(nc), grid.shape[1]))
    dth, dpts = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc, want_dpoints=True)
    dth, dpts = dth.cpu().numpy(), dpts.cpu().numpy()

    def loss_checker(th):
        return float((O.closed_form_nd(grid, O.theta_to_affine(B, th, nc, np.float64), nc) * gout).sum())

    def loss_kernel(th):
        A = O.theta_to_affine(B, th, nc, np.float64)
        return float((ops.forward_closed_form(dev(grid), dev(A), nc).cpu().numpy() * gout).sum())

    eps = 1e-6
    for k in range(2):                                  # checker, random directions
        dirn = rng.normal(size=theta.shape)
        fd = (loss_checker(theta + eps * dirn) - loss_checker(theta - eps * dirn)) / (2 * eps)
        an = float((dth * dirn).sum())
        print("%s: directional derivative %.9e, differenced checker %.9e" % (name, an, fd))
        assert abs(an - fd) < 2e-6 * max(1.0, abs(fd))
    fd = np.zeros_like(theta)                           # kernel forward, every component
    for t in range(theta.shape[0]):
        for k in range(theta.shape[1]):
            e = np.zeros_like(theta); e[t, k] = eps
            fd[t, k] = (loss_kernel(theta + e) - loss_kernel(theta - e)) / (2 * eps)
    print("%s: dtheta vs differenced forward: rel err %.2e" % (name, rel_err(dth, fd)))
    assert rel_err(dth, fd) < 2e-6
    # d/dpoints: shift all points (central differences of the kernel's forward, one coordinate at a time)
    fdp = np.zeros_like(dpts)
    h = 1e-7
    for j in range(len(nc)):
        sh = np.zeros((len(nc), 1)); sh[j] = h
        up = ops.forward_closed_form(dev(grid + sh), dev(As), nc).cpu().numpy()
        dn = ops.forward_closed_form(dev(grid - sh), dev(As), nc).cpu().numpy()
        fdp[:, j] = ((up - dn) / (2 * h) * gout).sum(axis=1)
    interior = ((grid > 1e-3) & (grid < 1 - 1e-3)).all(axis=0)
    print("%s: dpoints vs differenced forward: rel err %.2e" % (name, rel_err(dpts[:, :, interior], fdp[:, :, interior])))
    assert rel_err(dpts[:, :, interior], fdp[:, :, interior]) < 1e-5
    # the adjoint started from the forward's own output (what autograd does) instead of walking forward first
    x1 = ops.forward_closed_form(dev(grid), dev(As), nc)
    dth_x1, dpts_x1 = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc, want_dpoints=True, newpoints=x1)
    print("%s: dtheta from the saved output vs from the in-kernel forward walk: rel err %.2e" % (name, rel_err(dth_x1.cpu().numpy(), dth)))
    # (d/dpoints away from the domain boundary: at a vertex of the tessellation that the flow keeps fixed, the
    # derivative of the flow is one-sided and depends on which of the adjoining simplices the walk starts in)
    assert rel_err(dth_x1.cpu().numpy(), dth) < 1e-10
    assert rel_err(dpts_x1.cpu().numpy()[:, :, interior], dpts[:, :, interior]) < 1e-9
    dth32, _ = ops.backward_theta_closed_form(dev(grid, torch.float32), dev(As, torch.float32), dev(B, torch.float32),
                                              dev(gout, torch.float32), nc)
    print("%s: float32 dtheta vs float64: rel err %.2e" % (name, rel_err(dth32.cpu().numpy(), dth)))
    assert rel_err(dth32.cpu().numpy(), dth) < 1e-4


@pytest.mark.parametrize("nc", [[3, 3], [2, 2, 2]])
def test_nd_gradient_at_the_identity(nc):
    """theta = 0: the field vanishes, and the exact gradient G_c = int lambda [x;1]^T dt over the cell of
    each (resting) point equals the fixed-step adjoint's, which is exact for a zero field."""
    from libcpab_b200 import ops
    name = "d2_t3x3" if len(nc) == 2 else "d3_t2x2x2"
    g = load_golden(name)
    B = g["B"]
    grid = g["grid"].astype(np.float64)[:, ::5]
    As = np.zeros((2, O.n_cells(nc), len(nc), len(nc) + 1))
    rng = np.random.default_rng(4)
    gout = rng.normal(size=(2, len(nc), grid.shape[1]))
    dth, dpts = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc, want_dpoints=True)
    ref, _ = ops.backward_theta(dev(grid), dev(As), dev(B), dev(gout), nc, 50)
    assert float(ref.abs().max()) > 1e-3
    assert rel_err(dth.cpu().numpy(), ref.cpu().numpy()) < 1e-9
    assert np.allclose(dpts.cpu().numpy(), gout)


@pytest.mark.parametrize("nc,size,n_theta", [([10, 10], [256, 256], 16), ([4, 4, 4], [48, 48, 48], 4)])
def test_nd_api_switch_and_flow_properties(nc, size, n_theta):
    from libcpab_b200 import Cpab
    torch.manual_seed(0)
    T = Cpab(nc, backend="pytorch", device="gpu", volume_perservation=len(nc) == 2)
    T.params.closed_form = True
    theta = (0.5 * T.sample_transformation(n_theta)).requires_grad_(True)
    grid = T.uniform_meshgrid(size)
    out = T.transform_grid(grid, theta)
    assert tuple(out.shape) == (n_theta, len(nc), grid.shape[1]) and bool(torch.isfinite(out).all())
    assert float(out.min()) > -1e-6 and float(out.max()) < 1 + 1e-6          # zero boundary: the box maps to itself
    back = T.transform_grid(out.detach(), -theta.detach())                     # inverse flow
    print("inverse flow error %.2e" % float((back - grid[None]).abs().max()))
    assert float((back - grid[None]).abs().max()) < 2e-5
    out.square().sum().backward()
    assert bool(torch.isfinite(theta.grad).all()) and float(theta.grad.abs().max()) > 0
    T.params.closed_form = False
    fixed = T.transform_grid(grid, theta.detach())
    # (away from the upper faces: a coordinate of exactly 1 sends the reference's float32 cell search to the
    # wrong simplex of the last cube, cpab_ops.cpp:139-148 / SURVEY.md 7.3 -- the fixed-step kernels reproduce that)
    inner = (grid < 1).all(dim=0)
    diff = (fixed - out.detach())[:, :, inner].abs().max()
    print("fixed-step (50) vs hit-time: %.2e (all points: %.2e)" % (float(diff), float((fixed - out.detach()).abs().max())))
    assert float(diff) < 5e-3
    T.params.closed_form = True
    assert float((T.transform_grid(grid, T.identity(2)) - grid[None]).abs().max()) < 1e-6
    # transform_data goes through the same switch
    data = torch.rand(n_theta, 1, *size, device="cuda")
    img = T.transform_data(data, theta.detach(), size)
    assert tuple(img.shape) == tuple(data.shape) and bool(torch.isfinite(img).all())


@pytest.mark.parametrize("nc,size,n_theta", [([10, 10], [512, 512], 8), ([4, 4, 4], [80, 80, 80], 4)])
def test_nd_float32_agrees_with_float64_at_scale(nc, size, n_theta):
    """Millions of trajectories, theta ~ N(0, I): the float32 kernels against the float64 ones (which the small
    cases hold to the checker) -- the net for rare paths (vertices, grazing faces, dips between scan nodes)."""
    from libcpab_b200 import Cpab, ops
    torch.manual_seed(5)
    T = Cpab(nc, backend="pytorch", device="gpu")
    theta = T.sample_transformation(n_theta).double()
    grid = T.uniform_meshgrid(size).double()
    B = torch.as_tensor(np.asarray(T.params.basis), dtype=torch.float64, device="cuda")
    As = (B @ theta.T).T.reshape(n_theta, -1, len(nc), len(nc) + 1).contiguous()
    gout = torch.randn(n_theta, len(nc), grid.shape[1], dtype=torch.float64, device="cuda")
    x64 = ops.forward_closed_form(grid, As, nc)
    x32 = ops.forward_closed_form(grid.float(), As.float(), nc)
    err = (x32.double() - x64).abs().amax(dim=1)
    gain = flow_gain(As.cpu().numpy())
    print("nc %s: %d trajectories, float32 vs float64 max %.2e, 99.99th percentile %.2e (flow gain %.1f)"
          % (nc, err.numel(), float(err.max()), float(err.flatten().kthvalue(int(err.numel() * 0.9999)).values), gain))
    assert float(err.max()) < 1e-6 * gain + 5e-6
    d64, _ = ops.backward_theta_closed_form(grid, As, B, gout, nc, newpoints=x64)
    d32, _ = ops.backward_theta_closed_form(grid.float(), As.float(), B.float(), gout.float(), nc, newpoints=x32)
    e = theta_rel = ((d32.double() - d64).abs().amax(dim=1) / d64.abs().amax(dim=1))
    print("nc %s: dtheta float32 vs float64, per theta: max %.2e median %.2e" % (nc, float(e.max()), float(e.median())))
    assert float(e.max()) < 1e-3        # (the reference's own float32 and float64 gradients differ by 9e-4 on BASELINE configs[0])
    back = ops.forward_closed_form(x64, -As, nc)                      # the flow of -theta undoes it
    assert float((back - grid[None]).abs().max()) < 1e-10
    # the adjoint over all these trajectories against central differences of the float64 forward, random directions
    for k in range(2):
        dirn = torch.randn_like(theta)
        # (the flow is C1 in theta with a kink in its second derivative at every change of a cell sequence; summed over
        #  2e6 trajectories the central difference is only good to ~1e-5 relative in 2-D -- it moves by 1.3e-4 between
        #  eps = 1e-6 and 1e-5, and a Richardson step does not cure it; the 64-point cases above hold the adjoint to 1e-9)
        eps = 1e-6

        def loss(th):
            A = (B @ th.T).T.reshape(n_theta, -1, len(nc), len(nc) + 1).contiguous()
            return float((ops.forward_closed_form(grid, A, nc) * gout).sum())

        fd = (loss(theta + eps * dirn) - loss(theta - eps * dirn)) / (2 * eps)
        an = float((d64 * dirn).sum())
        print("nc %s: directional derivative over %d trajectories: adjoint %.10e, differenced forward %.10e"
              % (nc, err.numel(), an, fd))
        assert abs(an - fd) < 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("nc,size,n_theta", [([10, 10], [256, 256], 32), ([4, 4, 4], [64, 64, 64], 4)])
def test_nd_lane_utilisation_with_and_without_refill(nc, size, n_theta):
    """Divergence of the crossing loop, measured: share of the warps' loop slots that do work."""
    from libcpab_b200 import Cpab, _lib, ops
    torch.manual_seed(1)
    T = Cpab(nc, backend="pytorch", device="gpu")
    theta = T.sample_transformation(n_theta)
    grid = T.uniform_meshgrid(size)
    B = torch.as_tensor(np.asarray(T.params.basis), dtype=torch.float32, device="cuda")
    As = (B @ theta.T).T.reshape(n_theta, -1, len(nc), len(nc) + 1).contiguous()
    res = {}
    try:
        for refill in (0, 1):
            _lib.set_tuning("closed_refill", refill)
            out, util, per_traj = ops.closed_form_lane_stats(grid, As, nc)
            res[refill] = (out, util, per_traj)
    finally:
        _lib.set_tuning("closed_refill", 1)
    print("nc %s: lane utilisation %.3f without refill, %.3f with; %.2f sub-steps per trajectory"
          % (nc, res[0][1], res[1][1], res[1][2]))
    assert torch.equal(res[0][0], res[1][0])                 # the same trajectories either way
    assert res[1][1] > res[0][1] and res[1][1] > 0.85
    assert abs(res[0][2] - res[1][2]) < 1e-9


@pytest.mark.parametrize("nc", [[3, 3], [2, 2, 2]])
def test_nd_walk_terminates_on_wild_fields(nc):
    """Non-finite and absurdly large velocity fields: the sub-step loop has a hard trip-count bound, so the
    kernel returns (NaN / far-away points are fine, a hang is not)."""
    from libcpab_b200 import ops
    n = len(nc)
    rng = np.random.default_rng(9)
    grid = rng.uniform(0, 1, size=(n, 96)).astype(np.float32)
    gout = rng.normal(size=(4, n, 96)).astype(np.float32)
    As = rng.normal(size=(4, O.n_cells(nc), n, n + 1)).astype(np.float32)
    As[0] *= 1e6                      # 2e6 sub-steps would be needed: cut off at the bound
    As[1, 3] = np.nan
    As[2, 5, 0, 0] = np.inf
    B = rng.normal(size=(O.n_cells(nc) * n * (n + 1), 5)).astype(np.float32)
    out = ops.forward_closed_form(dev(grid), dev(As), nc)
    dth, _ = ops.backward_theta_closed_form(dev(grid), dev(As), dev(B), dev(gout), nc)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (4, n, 96) and tuple(dth.shape) == (4, 5)
    assert bool(torch.isfinite(out[3]).all()) and bool(torch.isfinite(dth[3]).all())      # the sane theta is untouched


def test_nd_sequential_chain_trains_every_warp_exactly():
    """CpabSequential of two 2-D hit-time flows with params.points_grad: the gradient reaches the first warp through
    lambda(0) of the second (dL/dpoints of the adjoint).  Both thetas against central differences of the chain, float64."""
    from libcpab_b200 import Cpab, CpabSequential
    g = load_golden("d2_t3x3")
    Ts = [Cpab([3, 3], backend="pytorch", device="gpu", basis=g["B"]) for _ in range(2)]
    for T in Ts:
        T.params.closed_form = True
        T.params.points_grad = True
    S = CpabSequential(*Ts)
    torch.manual_seed(2)
    thetas = [(0.6 * torch.as_tensor(g["theta"][k:k + 2], dtype=torch.float64)).cuda().requires_grad_(True) for k in (0, 2)]
    grid = torch.rand(2, 300, dtype=torch.float64, device="cuda") * 0.98 + 0.01
    R = torch.randn(2, 2, 300, dtype=torch.float64, device="cuda")

    def loss(ths):
        return (S.transform_grid(grid, ths) * R).sum()

    loss(thetas).backward()
    rng = np.random.default_rng(0)
    for w in range(2):
        assert thetas[w].grad is not None
        for _ in range(3):
            dirn = torch.as_tensor(rng.normal(size=tuple(thetas[w].shape)), device="cuda")
            eps = 1e-6
            plus = [t.detach().clone() for t in thetas]
            minus = [t.detach().clone() for t in thetas]
            plus[w] += eps * dirn
            minus[w] -= eps * dirn
            fd = float(loss(plus) - loss(minus)) / (2 * eps)
            an = float((thetas[w].grad * dirn).sum())
            print("warp %d: directional derivative %.9e, differenced chain %.9e" % (w, an, fd))
            assert abs(an - fd) < 1e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("name", ["d2_t10x10_vp", "d3_t2x2x2"])
def test_nd_unstaged_tables_give_the_same_result(name):
    """Tessellations whose velocity matrices do not fit shared memory read them through L1; forced here with
    the tuning key "closed_stage" = 0: identical trajectories and gradients."""
    from libcpab_b200 import _lib, ops
    g, nc, theta, As, grid = _nd_case(name, 400)
    rng = np.random.default_rng(8)
    gout = rng.normal(size=(theta.shape[0], len(nc), grid.shape[1]))
    res = []
    try:
        for stage in (1, 0):
            _lib.set_tuning("closed_stage", stage)
            x = ops.forward_closed_form(dev(grid, torch.float32), dev(As, torch.float32), nc)
            d, _ = ops.backward_theta_closed_form(dev(grid), dev(As), dev(g["B"]), dev(gout), nc)
            res.append((x, d))
    finally:
        _lib.set_tuning("closed_stage", 1)
    assert torch.equal(res[0][0], res[1][0])
    assert rel_err(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) < 1e-12      # (atomics: summation order)


def test_closed_form_rejects_bad_arguments():
    from libcpab_b200 import _lib, ops
    with pytest.raises(_lib.CpabError):
        ops.closed_form_lane_stats(torch.zeros(1, 8, device="cuda"), torch.zeros(1, 4, 1, 2, device="cuda"), [4])
