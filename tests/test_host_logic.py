"""Host-side logic that needs no GPU: tessellation / basis construction, the Cpab API surface and
its argument checks, theta-sharding and the gradient all-reduce (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------ tessellation
@pytest.mark.parametrize("name", golden_cases())
def test_basis_spans_the_reference_null_space(name, tmp_path):
    from libcpab_b200.tessellation import Tessellation
    g = load_golden(name)
    T = Tessellation(g["nc"].tolist(), zero_boundary=bool(g["zero_boundary"]),
                     volume_perservation=bool(g["volume_perservation"]), direc=str(tmp_path))
    B, Bref = T.B, g["B"]
    assert B.shape == Bref.shape
    assert np.abs(B.T @ B - np.eye(B.shape[1])).max() < 1e-12
    assert np.abs(Bref - B @ (B.T @ Bref)).max() < 1e-12          # same subspace
    assert np.abs(T.L @ Bref).max() < 1e-9                        # reference basis obeys OUR constraints
    # cached copy is read back unchanged
    T2 = Tessellation(g["nc"].tolist(), zero_boundary=bool(g["zero_boundary"]),
                      volume_perservation=bool(g["volume_perservation"]), direc=str(tmp_path))
    assert np.array_equal(T2.B, B)


def test_basis_fields_are_continuous_and_respect_the_boundary():
    from libcpab_b200.tessellation import Tessellation, cell_vertices
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    T = Tessellation([4, 3], zero_boundary=True, volume_perservation=True, direc="/tmp/_basis_t")
    theta = rng.standard_normal(T.B.shape[1])
    A = (T.B @ theta).reshape(-1, 2, 3)
    assert np.abs(A[:, 0, 0] + A[:, 1, 1]).max() < 1e-12          # zero trace
    # velocity is single valued at points sitting on shared edges
    pts = rng.uniform(0, 1, (2, 4000)).astype(np.float32)
    eps = 1e-4
    for shift in ([eps, 0], [0, eps]):
        a = O.findcellidx(pts, [4, 3])
        q = (pts + np.asarray(shift, np.float32)[:, None]).astype(np.float32)
        b = O.findcellidx(q, [4, 3])
        va = np.einsum("pij,jp->ip", A[a], np.vstack([pts, np.ones(4000)]))
        vb = np.einsum("pij,jp->ip", A[b], np.vstack([q, np.ones(4000)]))
        assert np.abs(va - vb).max() < 20 * eps * np.abs(A).max()
    # normal velocity vanishes on the boundary
    edge = np.vstack([np.zeros(50), np.linspace(0.01, 0.99, 50)]).astype(np.float32)
    ce = O.findcellidx(edge, [4, 3])
    v = np.einsum("pij,jp->ip", A[ce], np.vstack([edge, np.ones(50)]))
    assert np.abs(v[0]).max() < 1e-12
    assert cell_vertices([4, 3]).shape == (48, 3, 3)


# ------------------------------------------------------------------------------------ API surface
def test_cpab_constructor_and_checks_without_a_gpu():
    from libcpab_b200 import Cpab, CpabSequential
    T = Cpab([3, 3], backend="pytorch", device="gpu")
    assert T.get_theta_dim() == 34 and T.params.nC == 36 and T.params.D == 216
    assert T.params.nstepsolver == 50 and T.get_basis().shape == (216, 34)
    for bad in (dict(tess_size=[]), dict(tess_size=[1, 2, 3, 4]), dict(tess_size=[2.0]),
                dict(tess_size=[0]), dict(tess_size=[2], backend="jax"), dict(tess_size=[2], device="tpu"),
                dict(tess_size=[2], zero_boundary=1)):
        kw = dict(tess_size=[2], backend="pytorch", device="gpu")
        kw.update(bad)
        with pytest.raises(AssertionError):
            Cpab(**kw)
    for unsupported in (dict(backend="numpy", device="cpu"), dict(backend="pytorch", device="cpu"),
                        dict(backend="tensorflow", device="gpu")):
        with pytest.raises(NotImplementedError):
            Cpab([2], **unsupported)
    with pytest.raises(AssertionError):
        T.set_solver_params(nstepsolver=0)
    with pytest.raises(NotImplementedError):
        T.set_solver_params(numeric_grad=True)
    T.set_solver_params(nstepsolver=20)
    assert T.params.nstepsolver == 20
    # CPU tensors are rejected by the device check, like the reference (cpab.py:513-520)
    with pytest.raises(AssertionError):
        T.transform_grid(torch.zeros(2, 4), torch.zeros(1, 34))
    with pytest.raises(AssertionError):
        CpabSequential(T, Cpab([5], backend="pytorch", device="gpu"))   # mixed dimensionality
    S = CpabSequential(T, Cpab([2, 2], backend="pytorch", device="gpu"))
    assert S.get_theta_dim() == [34, S.cpab[1].params.d]
    with pytest.raises(AssertionError):
        S.transform_grid(torch.zeros(2, 4), [torch.zeros(1, 34)])       # one theta missing


def test_ops_refuse_cpu_tensors_loudly():
    from libcpab_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.forward(torch.zeros(2, 8), torch.zeros(1, 36, 2, 3), [3, 3], 50)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.interpolate_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 16), [4, 4])


def test_backend_helpers_match_the_reference_conventions():
    from libcpab_b200 import functions as F
    g = F.uniform_meshgrid(2, [0, 0], [1, 1], [4, 3], device="cpu")
    assert tuple(g.shape) == (2, 12)
    assert torch.equal(g[0, :4], torch.linspace(0, 1, 4)) and torch.equal(g[1, ::4], torch.linspace(0, 1, 3))
    assert torch.equal(F.identity(5, 2, epsilon=1e-6), torch.full((2, 5), 1e-6))
    assert tuple(F.sample_transformation(7, 3).shape) == (3, 7)
    assert F.backend_type() is torch.Tensor and F.check_device(torch.zeros(1), "cpu")
    m = torch.rand(5, 2)
    assert torch.allclose(F.pdist(m), torch.cdist(m, m) ** 2, atol=1e-5)


# ------------------------------------------------------------------------------------ distributed
def test_shard_bounds_cover_and_balance():
    from libcpab_b200.distributed import shard_bounds
    for n in (0, 1, 7, 64, 65, 65536):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libcpab_b200 import distributed as D
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "d2_t3x3.npz"))
    nc = g["nc"].tolist()
    theta = torch.from_numpy(g["theta"])                      # 4 thetas -> 2 per rank (world 2)
    local = D.shard(theta)
    lo, hi = D.shard_bounds(theta.shape[0])
    # stand-in for the GPU kernels in this CPU test: the oracle (allowed in tests only)
    As = O.theta_to_affine(g["B"], local.numpy(), nc)
    out_local = torch.from_numpy(O.forward(g["grid"], O.affine_to_trels(As), nc, 50))
    whole = D.gather_shards(out_local, theta.shape[0])
    # shared parameter: every rank contributes its shard's gradient, all-reduce sums them
    shared = torch.zeros(3, requires_grad=True)
    (shared * float(local.sum())).sum().backward()
    D.allreduce_grad_(shared)
    ok = torch.equal(whole[lo:hi], out_local) and whole.shape[0] == theta.shape[0]
    if rank == 0:
        torch.save({"whole": whole, "grad": shared.grad.clone(), "ok": ok}, out)
    dist.destroy_process_group()


def test_theta_sharding_and_gradient_allreduce_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    g = load_golden("d2_t3x3")
    nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"], nc)
    single = O.forward(g["grid"], O.affine_to_trels(As), nc, 50)
    assert res["ok"]
    assert np.array_equal(res["whole"].numpy(), single)        # sharded == unsharded, bitwise
    assert torch.allclose(res["grad"], torch.full((3,), float(g["theta"].sum())), rtol=1e-5)


# ----------------------------------------------------------------- point sharding (configs[3] split)
class _OracleCpab:
    """CPU stand-in with the three methods PointShardedCpab calls; backed by the oracle (tests only)."""

    def __init__(self, g):
        from oracle import oracle as O
        self.O, self.g, self.nc = O, g, g["nc"].tolist()

    def uniform_meshgrid(self, outsize):
        return torch.from_numpy(self.O.uniform_meshgrid(outsize))

    def transform_grid(self, grid, theta):
        As = self.O.theta_to_affine(self.g["B"], theta.numpy(), self.nc)
        return torch.from_numpy(self.O.forward(grid.numpy(), self.O.affine_to_trels(As), self.nc, 50))

    def interpolate(self, data, grid, outsize):
        return torch.from_numpy(self.O.interpolate(data.numpy(), grid.numpy(), outsize))


def test_point_shard_bounds_are_slabs_of_the_last_dimension():
    from libcpab_b200.distributed import PointShardedCpab
    for outsize in ([7], [5, 9], [4, 3, 10]):
        for ws in (1, 2, 3, 8):
            spans = [PointShardedCpab(None, r, ws).point_bounds(outsize) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == int(np.prod(outsize))
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            stride = int(np.prod(outsize[:-1]))
            assert all(lo % stride == 0 and hi % stride == 0 for lo, hi in spans)
            assert sum(PointShardedCpab(None, r, ws).local_outsize(outsize)[-1] for r in range(ws)) == outsize[-1]


def _point_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libcpab_b200.distributed import PointShardedCpab
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "d2_t3x3.npz"))
    T = _OracleCpab(g)
    P = PointShardedCpab(T)
    outsize = [12, 9]
    theta = torch.from_numpy(g["theta"])
    rng = np.random.default_rng(0)
    data = torch.from_numpy(rng.random((theta.shape[0], 2, 8, 7), dtype=np.float32))
    local = P.transform_data_local(data, theta, outsize)
    whole = P.gather_data(local, outsize)
    # the collective: partial theta-gradients (here: a stand-in that sums the local slab) are summed
    theta.grad = torch.full_like(theta, float(local.sum()))
    P.allreduce_theta_grad_(theta)
    if rank == 0:
        torch.save({"whole": whole, "grad": theta.grad.clone()}, out)
    dist.destroy_process_group()


def test_point_sharding_world2_gloo(tmp_path):
    """Two ranks, each a slab of the output grid: gathered slabs == the unsharded result (bitwise),
    the all-reduce sums the per-slab contributions."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "p0.pt")
    mp.spawn(_point_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    g = load_golden("d2_t3x3")
    T = _OracleCpab(g)
    outsize = [12, 9]
    theta = torch.from_numpy(g["theta"])
    rng = np.random.default_rng(0)
    data = torch.from_numpy(rng.random((theta.shape[0], 2, 8, 7), dtype=np.float32))
    full = T.interpolate(data, T.transform_grid(T.uniform_meshgrid(outsize), theta), outsize)
    assert torch.equal(res["whole"], full)
    assert torch.allclose(res["grad"], torch.full_like(theta, float(full.sum())), rtol=1e-5)
