"""The drop-in behind the reference's OWN class.

The unmodified reference package (baseline/_ref/libcpab, copied there by __graft_entry__.build();
git-ignored, travels to the GPU box) is imported with `cpab_gpu` -- the module its
libcpab/pytorch/transformer.py:49-55 JIT-builds from its CUDA sources -- replaced by
integration/cpab_b200.py, the ctypes stub of INTEGRATION.md section 1 on top of libcpab_b200.so.
Everything else is the reference's code: `Cpab.__init__` (cpab.py:60-122), `transform_grid`
(:257-278), `transform_data` (:304-331), `_CPABFunction_AnalyticGrad` (pytorch/transformer.py:136-202)
with its torch expm, and its torch interpolation.  Results are compared with the fixtures the same
reference produced on the CPU (tests/golden) and with libcpab_b200's own `Cpab`.
"""
import numpy as np
import pytest
import torch

from conftest import flow_gain, golden_cases, load_golden, rel_err
from oracle import oracle as O
from oracle import ref_package

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_package.available(), reason="baseline/_ref/libcpab not installed")]
TOL = 1e-5


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ref_T(g):
    ref = ref_package.import_with_b200_backend()
    T = ref.Cpab(g["nc"].tolist(), backend="pytorch", device="gpu", zero_boundary=bool(g["zero_boundary"]),
                 volume_perservation=bool(g["volume_perservation"]), override=False)
    T.params.basis = g["B"]          # the null-space basis is unique only up to a rotation: use the fixture's
    assert T.params.basis.shape == (T.params.D, T.params.d)
    return T


def interior(g):
    if len(g["nc"]) < 3:
        return np.ones(g["grid"].shape[1], dtype=bool)
    return ((g["grid"] > 0.02) & (g["grid"] < 0.98)).all(axis=0)


def test_the_reference_runs_on_the_stub():
    ref = ref_package.import_with_b200_backend()
    from libcpab.pytorch import transformer as rt
    assert rt._gpu_succes and rt.cpab_gpu.__name__ == "cpab_b200_stub"
    assert ref.Cpab.__module__ == "libcpab.cpab"


@pytest.mark.parametrize("name", golden_cases())
def test_reference_cpab_transform_grid_forward_and_backward(name):
    from libcpab.pytorch.expm import expm as ref_expm
    g = load_golden(name)
    nc = g["nc"].tolist()
    T = ref_T(g)
    theta = cuda(g["theta"]).requires_grad_(True)
    grid = cuda(g["grid"])
    out = T.transform_grid(grid, theta)                       # reference code path -> cpab_b200 stub -> k_forward
    assert tuple(out.shape) == g["grid_t"].shape and out.is_cuda
    got = out.detach().cpu().numpy()
    # (1) exactly: the stub's forward is the oracle's forward on the Trels the reference computed on this GPU
    p = T.params
    Bt = torch.Tensor(p.basis).cuda()
    As = torch.matmul(Bt, theta.detach().t()).t().reshape(len(g["theta"]) * p.nC, *p.Ashape)
    sq = torch.cat([As, torch.zeros(len(g["theta"]) * p.nC, 1, p.ndim + 1, device="cuda")], dim=1)
    Tr = ref_expm((1.0 / p.nstepsolver) * sq)[:, :p.ndim, :].reshape(len(g["theta"]), p.nC, *p.Ashape).contiguous()
    assert np.array_equal(got, O.forward(g["grid"], Tr.cpu().numpy(), nc, 50))
    # (2) end to end against what the reference produced on the CPU
    gain = flow_gain(g["As"])
    keep = interior(g)
    e_fwd = rel_err(got[:, :, keep], g["grid_t"][:, :, keep])
    (out * cuda(g["gout"])).sum().backward()                  # reference backward -> stub -> k_jacobian -> mul_/sum
    e_bwd = rel_err(theta.grad.cpu().numpy(), g["dtheta"])
    print("reference Cpab on libcpab_b200 %s: forward rel err %.3g, dtheta rel err %.3g (flow gain %.2f)"
          % (name, e_fwd, e_bwd, gain))
    assert e_fwd < 2 * TOL * gain
    assert e_bwd < (2 * TOL * gain if len(nc) < 3 else 5e-3)
    # (3) the same call through libcpab_b200's own Cpab (fast path: fused projection + expm, adjoint gradient)
    from libcpab_b200 import Cpab
    T2 = Cpab(nc, backend="pytorch", device="gpu", zero_boundary=bool(g["zero_boundary"]),
              volume_perservation=bool(g["volume_perservation"]), basis=g["B"])
    th2 = cuda(g["theta"]).requires_grad_(True)
    out2 = T2.transform_grid(grid, th2)
    (out2 * cuda(g["gout"])).sum().backward()
    assert rel_err(out2.detach().cpu().numpy()[:, :, keep], got[:, :, keep]) < 2 * TOL * gain
    assert rel_err(th2.grad.cpu().numpy(), theta.grad.cpu().numpy()) < (2 * TOL * gain if len(nc) < 3 else 5e-3)


@pytest.mark.parametrize("name", ["d1_t100", "d2_t3x3", "d2_t10x10_vp"])
def test_reference_cpab_transform_data(name):
    g = load_golden(name)
    T = ref_T(g)
    theta = cuda(g["theta"]).requires_grad_(True)
    data = cuda(g["data"]).requires_grad_(True)
    outsize = g["grid_n"].tolist()
    out = T.transform_data(data, theta, outsize=outsize)      # reference meshgrid + transformer + torch interpolation
    assert tuple(out.shape) == g["data_t"].shape
    err = np.abs(out.detach().cpu().numpy() - g["data_t"])
    (out * cuda(g["data_gout"])).sum().backward()
    e_th = rel_err(theta.grad.cpu().numpy(), g["data_dtheta"])
    e_dd = rel_err(data.grad.cpu().numpy(), g["data_ddata"])
    # the sampled image moves by (point error) x (steepest texel slope); the points agree to
    # 2e-5 x flow gain (previous test: torch's expm on the GPU rounds differently from the CPU's)
    slope = max(s - 1 for s in g["data"].shape[2:]) * np.abs(np.diff(g["data"], axis=-1)).max() * len(outsize)
    bound = 2 * TOL * flow_gain(g["As"]) * slope + 1e-6
    print("reference transform_data on libcpab_b200 %s: max |out - ref| %.3g (bound %.3g), median %.3g, dtheta %.3g, ddata %.3g"
          % (name, err.max(), bound, np.median(err), e_th, e_dd))
    mslope = max(s - 1 for s in g["data"].shape[2:]) * np.abs(np.diff(g["data"], axis=-1)).mean() * len(outsize)
    assert err.max() < bound and np.median(err) < 2 * TOL * flow_gain(g["As"]) * mslope + 1e-6
    assert e_th < 2e-3 and e_dd < 2e-3


def test_reference_aligner_on_the_stub():
    """The reference's own `CpabAligner` (libcpab/alignment.py:60-87) driving the reference's own `Cpab`
    on libcpab_b200.so: five Adam steps, against what the same code reached on the CPU."""
    ref = ref_package.import_with_b200_backend()
    g = load_golden("aux_align_2d")
    T = ref.Cpab(g["nc"].tolist(), backend="pytorch", device="gpu", zero_boundary=True, volume_perservation=False, override=False)
    T.params.basis = g["B"]
    A = ref.CpabAligner(T)
    theta = A.alignment_by_gradient(cuda(g["x1"]), cuda(g["x2"]), maxiter=5, lr=1e-2)
    err = float(np.abs(theta.detach().cpu().numpy() - g["theta_out"]).max())
    print("reference CpabAligner on libcpab_b200: max |theta - CPU theta| %.3g (|theta| ~ %.3g)" % (err, float(np.abs(g["theta_out"]).max())))
    # (Adam divides by the running gradient magnitude: last-bit differences of the first gradients --
    #  torch's own norm / interpolation on the GPU vs on the CPU included -- show at the 1e-4 level;
    #  the reference's own aligner on the GPU differs from itself on the CPU by 7e-5, this mirror by 4e-5)
    assert err < 2e-4
