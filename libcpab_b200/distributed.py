"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on B200 boxes,
gloo in the CPU tests).

The CPAB path shards naturally: every (theta, point) pair is independent in the
forward pass and the backward pass only couples the points of one theta.  So the theta batch --
and the matching rows of the data -- is split contiguously over ranks, every rank runs the
unchanged single-GPU kernels on its shard, and NO collective is issued on the forward path.
A collective is needed only where a parameter is SHARED across shards (alignment / training
mode): then the per-rank gradient of that parameter is summed with one all-reduce, enqueued on
the compute stream right after the gradient epilogue.

Secondary split, for few thetas on a huge grid (BASELINE configs[3], 16 thetas x 128^3): the POINTS
of the one problem are split over the ranks (`PointShardedCpab`).  Every rank integrates and
samples its slab for all thetas; the theta-gradient is a sum over points, so one all-reduce of
dtheta [n_theta, d] completes it (14 KB for configs[3]; all-reducing the per-cell G [n_theta, D]
instead would move 245 KB and repeat the G.B epilogue on every rank -- SURVEY.md 8-e).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int = None, world_size: int = None):
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (first n % ws ranks get one more)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int = None, world_size: int = None) -> torch.Tensor:
    """This rank's rows of a batch-major tensor (thetas, data, per-theta grids)."""
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def allreduce_grad_(param: torch.Tensor, average: bool = False) -> torch.Tensor:
    """Sum (or average) `param.grad` over ranks in place: the only collective of the path."""
    if param.grad is None:
        return param
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(param.grad, op=dist.ReduceOp.SUM)
        if average:
            param.grad.div_(ws)
    return param


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_shards(t: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenate theta-shards from every rank (verification / host-side collection only; the
    compute path never needs it).  Ragged shards are padded to the largest one."""
    rank, ws = world()
    if ws == 1:
        return t
    sizes = [shard_bounds(total, r, ws) for r in range(ws)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


class ShardedCpab:
    """Runs a Cpab (or CpabSequential) over this rank's shard of a theta/data batch.

    `transform_data(data, theta, outsize)` takes the FULL batch on every rank (or tensors that
    are already local when `presharded=True`), computes only the local rows and returns them;
    gradients of per-sample thetas stay local, gradients of shared leaves are summed by
    `allreduce_grad_`.
    """

    def __init__(self, T, presharded: bool = False):
        self.T = T
        self.presharded = presharded

    def _local(self, x):
        if self.presharded:
            return x
        if isinstance(x, (list, tuple)):
            return [shard(v) for v in x]
        return shard(x)

    def transform_grid(self, grid, theta):
        g = self._local(grid) if (hasattr(grid, "dim") and grid.dim() == 3) else grid
        return self.T.transform_grid(g, self._local(theta))

    def transform_data(self, data, theta, outsize):
        return self.T.transform_data(self._local(data), self._local(theta), outsize)


class PointShardedCpab:
    """Point-sharded `transform_grid` / `transform_data` of ONE problem over the ranks.

    The output grid `uniform_meshgrid(outsize)` is ordered first-coordinate-fastest, so a slab of
    the LAST output dimension is a contiguous range of points, and the matching slab of the
    sampled output `[N, C, ..., outsize[-1]]` is what `interpolate` produces for the local output
    size `[..., slab]`.  No collective on the forward path; `allreduce_theta_grad_` sums the
    partial theta-gradients after backward (the only collective).  `rank` / `world_size` default
    to the process group's; passing them explicitly lets one process play every rank (tests).
    """

    def __init__(self, T, rank: int = None, world_size: int = None):
        self.T = T
        r, ws = world()
        self.rank = r if rank is None else int(rank)
        self.world_size = ws if world_size is None else int(world_size)

    def slab_bounds(self, outsize):
        """[lo, hi) of the last output dimension owned by this rank."""
        return shard_bounds(int(outsize[-1]), self.rank, self.world_size)

    def point_bounds(self, outsize):
        """[lo, hi) of the meshgrid points owned by this rank."""
        stride = 1
        for v in outsize[:-1]:
            stride *= int(v)
        lo, hi = self.slab_bounds(outsize)
        return lo * stride, hi * stride

    def local_outsize(self, outsize):
        lo, hi = self.slab_bounds(outsize)
        return [int(v) for v in outsize[:-1]] + [hi - lo]

    def local_grid(self, outsize):
        lo, hi = self.point_bounds(outsize)
        return self.T.uniform_meshgrid(outsize)[:, lo:hi].contiguous()

    def transform_grid_local(self, theta, outsize):
        """[n_theta, ndim, local points]: this rank's slab of transform_grid(uniform_meshgrid)."""
        return self.T.transform_grid(self.local_grid(outsize), theta)

    def transform_data_local(self, data, theta, outsize):
        """[N, C, *outsize[:-1], slab]: this rank's slab of Cpab.transform_data(data, theta, outsize)."""
        grid_t = self.transform_grid_local(theta, outsize)
        return self.T.interpolate(data, grid_t, self.local_outsize(outsize))

    def allreduce_theta_grad_(self, theta: torch.Tensor) -> torch.Tensor:
        """Sum the partial dL/dtheta [n_theta, d] of the point shards (in place)."""
        if theta.grad is not None and self.world_size > 1 and dist.is_available() and dist.is_initialized():
            dist.all_reduce(theta.grad, op=dist.ReduceOp.SUM)
        return theta

    def gather_data(self, out_local: torch.Tensor, outsize) -> torch.Tensor:
        """Concatenate the slabs of every rank along the last dimension (verification only)."""
        if self.world_size == 1 or not (dist.is_available() and dist.is_initialized()):
            return out_local
        sizes = [shard_bounds(int(outsize[-1]), r, self.world_size) for r in range(self.world_size)]
        biggest = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros(tuple(out_local.shape[:-1]) + (biggest,), dtype=out_local.dtype, device=out_local.device)
        pad[..., :out_local.shape[-1]] = out_local
        parts = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(parts, pad)
        return torch.cat([p[..., :hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=-1)
