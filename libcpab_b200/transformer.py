"""`transformer(grid, theta, params)` -- the integrator behind Cpab.transform_grid.

Mirrors the role of libcpab/pytorch/transformer.py (CPAB_transformer ->
_CPABFunction_AnalyticGrad, :72-202) with the native calls replaced by the C ABI:

    forward : theta --cpab_b200_theta_to_trels--> (As, Trels) --cpab_b200_forward--> points
    backward: cpab_b200_backward_theta (adjoint sweep + G.B epilogue) -> dL/dtheta

Differences a caller can observe, all deliberate:
* the basis is uploaded to the device once per Cpab instance (the reference re-uploads it on
  every call, transformer.py:146);
* there is no slow path and no numeric-gradient path: `use_slow=True` / `numeric_grad=True`
  raise instead of silently changing algorithm (north_star: no fallback);
* the gradient w.r.t. `points` is None by default, exactly like the reference
  (transformer.py:202), so CpabSequential trains only its last warp unless
  `params.points_grad = True` is set, in which case the adjoint's lambda_0 is returned.
"""
from __future__ import annotations

import torch

from . import ops


class _BasisCache:
    """Device copies of B [D,d] and B^T [d,D] per (device, dtype), built lazily."""

    def __init__(self, basis):
        self._host = basis
        self._dev = {}

    def get(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._dev:
            b = torch.as_tensor(self._host, dtype=torch.float64).to(dtype).to(device).contiguous()
            self._dev[key] = (b, b.t().contiguous())
        return self._dev[key]


def _basis(params, device, dtype):
    cache = getattr(params, "_basis_cache", None)
    if cache is None or cache._host is not params.basis:
        cache = _BasisCache(params.basis)
        params._basis_cache = cache
    return cache.get(device, dtype)


class _CpabFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, theta, params):
        B, Bt = _basis(params, theta.device, theta.dtype)
        As, trels = ops.theta_to_trels(theta, Bt, params.nc, params.nstepsolver)
        ctx.closed_form = bool(getattr(params, "closed_form", False))
        if ctx.closed_form:      # opt-in extension (1-D): exact hit-time integration, no step count
            newpoints = ops.forward_closed_form(points, As, params.nc)
        else:
            newpoints = ops.forward(points, trels, params.nc, params.nstepsolver,
                                    fast_math=bool(getattr(params, "fast_math", False)))
        if ctx.closed_form and params.ndim > 1:       # the hit-time adjoint walks back from the output
            ctx.save_for_backward(points, As, B, newpoints)
        else:
            ctx.save_for_backward(points, As, B)
        ctx.params = params
        ctx.points_need_grad = points.requires_grad and bool(getattr(params, "points_grad", False))
        return newpoints

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        points, As, B = ctx.saved_tensors[:3]
        p = ctx.params
        if ctx.closed_form:
            x1 = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
            dtheta, dpoints = ops.backward_theta_closed_form(points, As, B, grad.contiguous(), p.nc,
                                                             want_dpoints=ctx.points_need_grad, newpoints=x1)
        else:
            dtheta, dpoints = ops.backward_theta(points, As, B, grad.contiguous(), p.nc, p.nstepsolver,
                                                 want_dpoints=ctx.points_need_grad,
                                                 fast_grad=bool(getattr(p, "fast_grad", False)))
        if dpoints is not None and points.dim() == 2:
            dpoints = dpoints.sum(dim=0)          # one grid shared by every theta
        return dpoints, dtheta, None


class _TransformDataFunction(torch.autograd.Function):
    """Cpab.transform_data as ONE forward and ONE backward kernel (plus the tiny theta->Trels and
    G.B kernels): the sampler runs as the epilogue of the integration kernel and its VJP as the
    prologue of the adjoint kernel.  Bit-identical to transform_grid followed by interpolate."""

    @staticmethod
    def forward(ctx, data, theta, grid, params, outsize):
        B, Bt = _basis(params, theta.device, theta.dtype)
        As, trels = ops.theta_to_trels(theta, Bt, params.nc, params.nstepsolver)
        out, grid_t = ops.transform_data_forward(grid, trels, data, params.nc, params.nstepsolver, outsize,
                                                 fast_math=bool(getattr(params, "fast_math", False)))
        ctx.save_for_backward(data, grid, grid_t, As, B)
        ctx.params = params
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        data, grid, grid_t, As, B = ctx.saved_tensors
        p = ctx.params
        grad = grad.contiguous()
        dtheta = ddata = None
        if ctx.needs_input_grad[1]:
            dtheta = ops.transform_data_backward(grid, As, B, data, grid_t, grad, p.nc, p.nstepsolver,
                                                 fast_grad=bool(getattr(p, "fast_grad", False)))
        if ctx.needs_input_grad[0]:
            _, ddata = ops.interpolate_backward(data, grid_t, grad, want_dgrid=False, want_ddata=True)
        return ddata, dtheta, None, None, None


def fused_transform_data(data, theta, grid, params, outsize):
    if not (data.is_cuda and theta.is_cuda and grid.is_cuda):
        raise RuntimeError("libcpab_b200 runs on CUDA tensors only (backend='pytorch', device='gpu')")
    if getattr(params, "use_slow", False) or getattr(params, "numeric_grad", False):
        raise NotImplementedError("libcpab_b200 implements the fast analytic path only")
    return _TransformDataFunction.apply(data, theta, grid, params, tuple(int(v) for v in outsize))


def CPAB_transformer(points, theta, params):
    if getattr(params, "use_slow", False):
        raise NotImplementedError("libcpab_b200 has no slow (pure python) integrator")
    if getattr(params, "numeric_grad", False):
        raise NotImplementedError("libcpab_b200 computes the analytic gradient only")
    if not (points.is_cuda and theta.is_cuda):
        raise RuntimeError("libcpab_b200 runs on CUDA tensors only (backend='pytorch', device='gpu')")
    if points.dtype != theta.dtype:
        raise TypeError("grid and theta must have the same dtype")
    return _CpabFunction.apply(points, theta, params)
