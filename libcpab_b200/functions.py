"""Backend module with the names libcpab.cpab.Cpab calls on `self.backend`
(libcpab/pytorch/functions.py).  Everything here is host-side torch plumbing except
`transformer`, `interpolate` and `findcellidx`, which go to the CUDA library.
"""
from __future__ import annotations

import torch

from . import ops
from .interpolation import interpolate            # noqa: F401  (part of the backend interface)
from .transformer import CPAB_transformer as transformer   # noqa: F401


def assert_version():
    major, minor = (int(v) for v in torch.__version__.split(".")[:2])
    assert (major, minor) >= (1, 0), "pytorch 1.0.0 or newer is required"


def _device(device):
    """'gpu' means the CURRENT CUDA device, resolved to a concrete index (the reference creates its
    tensors on the current device at every call; a bare torch.device('cuda') used as a cache key
    would pin whatever device was current at first use)."""
    if device is None:
        return None
    if not isinstance(device, torch.device):
        device = torch.device("cuda") if device in ("gpu", "cuda") else torch.device(device)
    if device.type == "cuda" and device.index is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def to(x, dtype=torch.float32, device=None):
    if isinstance(x, torch.Tensor):
        return x.to(dtype=dtype, device=_device(device))
    return torch.tensor(x, dtype=dtype, device=_device(device))


def tonumpy(x):
    return x.detach().cpu().numpy()


def check_device(x, device_name):
    return x.is_cuda == (device_name == "gpu")


def backend_type():
    return torch.Tensor


def pdist(mat):
    norm = torch.sum(mat * mat, 1).reshape(-1, 1)
    return norm - 2 * mat.mm(mat.t()) + norm.t()


def norm(x):
    return torch.norm(x)


def matmul(x, y):
    return torch.matmul(x, y)


def transpose(x):
    return x.t()


def exp(x):
    return torch.exp(x)


def zeros(*s, device=None):
    return torch.zeros(*s, device=_device(device))


def ones(*s, device=None):
    return torch.ones(*s, device=_device(device))


def arange(x):
    return torch.arange(x)


def repeat(x, reps):
    return x.repeat(reps)


def batch_repeat(x, reps):
    return x.repeat(reps, *(x.dim() * [1]))


def maximum(x):
    return x.max()


def sample_transformation(d, n_sample=1, mean=None, cov=None, device="cpu"):
    dev = _device(device)
    mean = torch.zeros(d, dtype=torch.float32, device=dev) if mean is None else mean
    cov = torch.eye(d, dtype=torch.float32, device=dev) if cov is None else cov
    try:
        dist = torch.distributions.MultivariateNormal(mean, cov)
        return dist.sample((n_sample,)).to(dev)
    except (ValueError, RuntimeError):
        # A covariance that is positive SEMI-definite up to rounding (the smooth prior of
        # Cpab.sample_transformation_with_prior in float32) has no Cholesky factor: sample through its
        # eigen-decomposition instead, as numpy.random.multivariate_normal -- the reference's numpy
        # backend, libcpab/numpy/functions.py:86-90 -- does.
        w, V = torch.linalg.eigh(0.5 * (cov + cov.t()).double())
        z = torch.randn(n_sample, d, dtype=torch.float64, device=cov.device)
        out = mean.double() + (z * w.clamp_min(0).sqrt()) @ V.t()
        return out.to(torch.float32).to(dev)


def identity(d, n_sample=1, epsilon=0, device="cpu"):
    assert epsilon >= 0, "epsilon need to be larger than 0"
    return torch.zeros(n_sample, d, dtype=torch.float32, device=_device(device)) + epsilon


_MESHGRID_CACHE = {}


def uniform_meshgrid(ndim, domain_min, domain_max, n_points, device="cpu", _share=False):
    """[ndim, nP] grid, first coordinate fastest (libcpab/pytorch/functions.py:102-108).

    The 1-D linspaces are evaluated on the host and uploaded (a few KB): torch's CPU and CUDA
    linspace kernels differ in the last bit, and a CPU-generated grid keeps the input
    bit-identical to what the reference's CPU path integrates.  The result is cached per
    (domain, size, device): the upload is a synchronous pageable copy that would otherwise stall
    the launch queue at the start of every transform_data call.  A fresh copy is returned, so
    callers may modify it.
    """
    dev = _device(device)
    key = (ndim, tuple(float(v) for v in domain_min), tuple(float(v) for v in domain_max),
           tuple(int(v) for v in n_points), str(dev))
    grid = _MESHGRID_CACHE.get(key)
    if grid is None:
        lin = [torch.linspace(domain_min[i], domain_max[i], n_points[i]).to(dev) for i in range(ndim)]
        mesh = torch.meshgrid(lin[::-1], indexing="ij")
        grid = torch.cat([g.reshape(1, -1) for g in mesh[::-1]], dim=0).contiguous()
        if len(_MESHGRID_CACHE) >= 16:
            _MESHGRID_CACHE.pop(next(iter(_MESHGRID_CACHE)))
        _MESHGRID_CACHE[key] = grid
    return grid if _share else grid.clone()     # _share: internal read-only use (transform_data)


def findcellidx(ndim, grid, nc):
    """Cell index of every grid point, int32 [nP] -- the integrators' own (C++-exact) search."""
    assert grid.shape[0] == ndim
    return ops.findcellidx(grid, nc)


def calc_vectorfield(grid, theta, params):
    """Velocity at every grid point for one theta (libcpab/pytorch/functions.py:111-129)."""
    B = to(params.basis, dtype=theta.dtype, device=theta.device)
    As = torch.matmul(B, theta.flatten()).reshape(params.nC, *params.Ashape)
    idx = findcellidx(params.ndim, grid, params.nc).long()
    homog = torch.cat((grid, torch.ones(1, grid.shape[1], device=grid.device, dtype=grid.dtype)), 0)
    return torch.einsum("pij,jp->ip", As[idx], homog)
