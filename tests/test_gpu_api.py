"""End-to-end parity through the reference-facing API (Cpab / CpabSequential / CpabAligner with
backend='pytorch', device='gpu') against fixtures produced by the reference's own API on the CPU.

Here theta is the input, so the GPU's own projection + expm are in the loop.  The reference's
float32 Pade differs from an exactly rounded exponential by ~1e-7 per entry, which 50 chained
steps amplify to a few 1e-6 (the reference's own fp32-vs-fp64 gap is 8e-6), so the
end-to-end bar is north_star's 1e-5 relative, not bit equality.
"""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_T(g):
    from libcpab_b200 import Cpab
    return Cpab(g["nc"].tolist(), backend="pytorch", device="gpu",
                zero_boundary=bool(g["zero_boundary"]),
                volume_perservation=bool(g["volume_perservation"]), basis=g["B"])


@pytest.mark.parametrize("name", golden_cases())
def test_transform_grid_and_theta_gradient(name):
    g = load_golden(name)
    T = make_T(g)
    theta = cuda(g["theta"]).requires_grad_(True)
    out = T.transform_grid(cuda(g["grid"]), theta)
    assert tuple(out.shape) == g["grid_t"].shape
    assert rel_err(out.detach().cpu().numpy(), g["grid_t"]) < TOL
    (out * cuda(g["gout"])).sum().backward()
    assert rel_err(theta.grad.cpu().numpy(), g["dtheta"]) < TOL


@pytest.mark.parametrize("name", [n for n in golden_cases() if "data" in load_golden(n).files])
def test_transform_data_forward_backward(name):
    g = load_golden(name)
    T = make_T(g)
    theta = cuda(g["theta"]).requires_grad_(True)
    data = cuda(g["data"]).requires_grad_(True)
    out = T.transform_data(data, theta, outsize=g["grid_n"].tolist())
    assert tuple(out.shape) == g["data_t"].shape
    # interpolation multiplies point errors by the local image slope * (size-1)
    scale = max(g["data"].shape[2:])
    assert rel_err(out.detach().cpu().numpy(), g["data_t"]) < TOL * scale
    (out * cuda(g["data_gout"])).sum().backward()
    assert rel_err(theta.grad.cpu().numpy(), g["data_dtheta"]) < 5e-4      # see note below
    assert rel_err(data.grad.cpu().numpy(), g["data_ddata"]) < TOL * scale
    # note: d(out)/d(grid) is piecewise constant in the sample position; a point that falls on
    # the other side of a texel boundary (|dx| ~ 1e-6) switches to the neighbouring slope, so the
    # theta-gradient through transform_data is compared at the looser, texel-aware bar


def test_interpolate_api_matches_reference():
    g = load_golden("d2_t3x3")
    T = make_T(g)
    grid = cuda(g["grid_t"]).requires_grad_(True)
    out = T.interpolate(cuda(g["data"]), grid, g["grid_n"].tolist())
    assert np.array_equal(out.detach().cpu().numpy(), g["interp_out"])
    (out * cuda(g["data_gout"])).sum().backward()
    assert rel_err(grid.grad.cpu().numpy(), g["interp_dgrid"]) < TOL


def test_sequential_matches_reference_including_missing_gradients():
    from libcpab_b200 import Cpab, CpabSequential
    g = load_golden("seq_1d20x3")
    Ts = [Cpab([20], backend="pytorch", device="gpu", basis=g["B"]) for _ in range(3)]
    S = CpabSequential(*Ts)
    thetas = [cuda(g[f"theta{i}"]).requires_grad_(True) for i in range(3)]
    out = S.transform_data(cuda(g["data"]), thetas, outsize=[128])
    assert rel_err(out.detach().cpu().numpy(), g["data_t"]) < TOL * 96
    grids = S.transform_grid(S.uniform_meshgrid([128]), thetas, output_all=True)
    for i in range(3):
        assert rel_err(grids[i].detach().cpu().numpy(), g["grids"][i]) < TOL
    (out * cuda(g["data_gout"])).sum().backward()
    # the reference hands no gradient to `points`, so only the last warp's theta is trained
    for i in range(2):
        assert thetas[i].grad is None or float(thetas[i].grad.abs().max()) == 0.0
    assert rel_err(thetas[2].grad.cpu().numpy(), g["dtheta2"]) < 5e-4


def test_sequential_points_grad_extension_trains_every_warp():
    from libcpab_b200 import Cpab, CpabSequential
    g = load_golden("seq_1d20x3")
    Ts = [Cpab([20], backend="pytorch", device="gpu", basis=g["B"]) for _ in range(3)]
    for T in Ts:
        T.params.points_grad = True
    S = CpabSequential(*Ts)
    thetas = [cuda(g[f"theta{i}"]).requires_grad_(True) for i in range(3)]
    S.transform_data(cuda(g["data"]), thetas, outsize=[128]).square().sum().backward()
    assert all(t.grad is not None and float(t.grad.abs().max()) > 0 for t in thetas)


def test_aligner_reduces_the_loss():
    from libcpab_b200 import Cpab, CpabAligner
    torch.manual_seed(0)
    T = Cpab([8], backend="pytorch", device="gpu")
    x = torch.linspace(0, 6.28, 200, device="cuda")
    x1 = torch.sin(x)[None, None]
    x2 = T.transform_data(x1, 0.3 * T.sample_transformation(1), outsize=(200,))
    A = CpabAligner(T)
    A.alignment_by_gradient(x1, x2, maxiter=30, lr=5e-2)
    assert float(A.losses[-1]) < 0.7 * float(A.losses[0])


def test_api_argument_checks_mirror_the_reference():
    from libcpab_b200 import Cpab
    T = Cpab([3, 3], backend="pytorch", device="gpu")
    grid = T.uniform_meshgrid([8, 8])
    theta = T.sample_transformation(2)
    with pytest.raises(AssertionError):
        T.transform_grid(grid.cpu(), theta)                       # wrong device
    with pytest.raises(AssertionError):
        T.transform_grid(grid.cpu().numpy(), theta)               # wrong type
    with pytest.raises(AssertionError):
        T.transform_grid(grid[None].repeat(3, 1, 1), theta)       # batch mismatch
    with pytest.raises(NotImplementedError):
        T.set_solver_params(use_slow=True)
    assert T.get_theta_dim() == T.params.d == theta.shape[1]
    idx = T.findcellidx(grid)
    assert idx.dtype == torch.int32 and int(idx.max()) < T.params.nC
    v = T.calc_vectorfield(grid, theta[:1])
    assert tuple(v.shape) == (2, 64)
