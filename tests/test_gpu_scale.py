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
    # oracle (the reference's float32 arithmetic) on a sub-sample of points, first thetas: every
    # theta within 1e-5 (the default mode follows the reference's cell sequences)
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(nP, size=min(512, nP), replace=False))
    Bs = np.ascontiguousarray(B.cpu().numpy().T.reshape(B.shape[1], -1, len(tess), len(tess) + 1))
    n_chk = min(8, n_theta)
    ref = O.theta_grad(grid[:, sel].cpu().numpy(), As[:n_chk].cpu().numpy(), Bs,
                       g1[:n_chk][:, :, sel].cpu().numpy(), tess, 50, threads=8)
    got, _ = ops.backward_theta(grid[:, sel].contiguous(), As[:n_chk], B, g1[:n_chk][:, :, sel].contiguous(), tess, 50)
    assert_grad_parity(got.cpu().numpy(), ref, 1e-5, what=str(tess))


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


@pytest.mark.parametrize("tess,n_theta,outsize,ws", [([4, 4, 4], 3, [24, 20, 22], 4), ([3, 3], 5, [40, 33], 3), ([50], 6, [1000], 8)])
def test_point_sharded_equals_unsharded(tess, n_theta, outsize, ws):
    """configs[3]-style split of the POINTS over `ws` ranks, played by one process: the slabs
    concatenate to the unsharded output bit for bit and the partial theta-gradients sum to the
    unsharded gradient (1e-6), which is what the NCCL all-reduce of dtheta delivers."""
    from libcpab_b200.distributed import PointShardedCpab
    torch.manual_seed(5)
    T = _T(tess)
    theta = T.sample_transformation(n_theta)
    data = torch.rand(n_theta, 2, *[s + 3 for s in outsize], device="cuda")
    Rfull = torch.randn(n_theta, 2, *outsize, device="cuda")
    T.params.fused_transform_data = False
    th = theta.clone().requires_grad_(True)
    full = T.transform_data(data, th, outsize)
    (full * Rfull).sum().backward()
    parts, gsum = [], torch.zeros_like(theta)
    for r in range(ws):
        P = PointShardedCpab(T, r, ws)
        lo, hi = P.slab_bounds(outsize)
        t = theta.clone().requires_grad_(True)
        out = P.transform_data_local(data, t, outsize)
        parts.append(out)
        if hi > lo:
            (out * Rfull[..., lo:hi]).sum().backward()
            gsum += t.grad
    assert bool((torch.cat(parts, dim=-1) == full).all())
    # (float32 sums in a different order: per-slab atomics, then the sum over slabs)
    assert rel_err(gsum.cpu().numpy(), th.grad.cpu().numpy()) < 3e-6


def test_point_sharded_nccl_two_gpus(tmp_path):
    """The same through torch.distributed/NCCL on two GPUs (skipped on a one-GPU box)."""
    import subprocess, sys, os
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ps.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from libcpab_b200 import Cpab
from libcpab_b200.distributed import PointShardedCpab
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
torch.manual_seed(5)
T = Cpab([4, 4, 4], backend="pytorch", device="gpu")
T.params.fused_transform_data = False
outsize = [24, 20, 22]
theta = T.sample_transformation(3)
data = torch.rand(3, 1, 27, 23, 25, device="cuda")
R = torch.randn(3, 1, *outsize, device="cuda")
th = theta.clone().requires_grad_(True)
(T.transform_data(data, th, outsize) * R).sum().backward()
P = PointShardedCpab(T)
lo, hi = P.slab_bounds(outsize)
t = theta.clone().requires_grad_(True)
out = P.transform_data_local(data, t, outsize)
(out * R[..., lo:hi]).sum().backward()
P.allreduce_theta_grad_(t)
err = float((t.grad - th.grad).abs().max() / th.grad.abs().max())
whole = P.gather_data(out, outsize)
ok = bool((whole == T.transform_data(data, theta, outsize)).all())
if rank == 0:
    print("RESULT", err, ok)
dist.destroy_process_group()
''' % root)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert float(line[1]) < 3e-6 and line[2] == "True"


def test_alignment_allreduce_nccl_two_gpus(tmp_path):
    """Alignment / training mode on two GPUs: every rank transforms its shard of the series with its
    own thetas and pulls it towards ONE shared template; the NCCL all-reduce of the template gradient
    (the only collective of the path) must give what a single process gets on the whole batch, and
    the per-series theta gradients must equal the unsharded ones (skipped on a one-GPU box)."""
    import subprocess, sys, os
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "al.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from libcpab_b200 import Cpab, CpabSequential
from libcpab_b200.distributed import shard, allreduce_grad_
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
torch.manual_seed(9)                                   # same global problem on every rank
Ts = [Cpab([20], backend="pytorch", device="gpu") for _ in range(2)]
for T in Ts:
    T.params.points_grad = True
S = CpabSequential(*Ts)
n = 64
thetas = [0.5 * T.sample_transformation(n) for T in Ts]
data = torch.rand(n, 1, 200, device="cuda")
template0 = torch.rand(1, 1, 160, device="cuda")

def step(th_list, series):
    template = template0.clone().requires_grad_(True)
    ths = [t.clone().requires_grad_(True) for t in th_list]
    out = S.transform_data(series, ths, outsize=[160])
    (out - template).square().sum().backward()
    return template, ths

t_full, th_full = step(thetas, data)                    # the whole batch in one process
t_loc, th_loc = step([shard(t) for t in thetas], shard(data))
allreduce_grad_(t_loc)                                  # NCCL all-reduce of the shared parameter's gradient
e_t = float((t_loc.grad - t_full.grad).abs().max() / t_full.grad.abs().max())
e_th = max(float((a.grad - shard(b.grad)).abs().max() / b.grad.abs().max()) for a, b in zip(th_loc, th_full))
worst = torch.tensor([e_t, e_th], device="cuda")
dist.all_reduce(worst, op=dist.ReduceOp.MAX)
if rank == 0:
    print("RESULT", float(worst[0]), float(worst[1]))
dist.destroy_process_group()
''' % root)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29535", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    print("alignment on 2 GPUs: template gradient rel diff %s, per-series theta gradients rel diff %s" % (line[1], line[2]))
    assert float(line[1]) < 3e-6 and float(line[2]) < 1e-6
