"""`interpolate(ndim, data, grid, outsize)` -- linear / bilinear / trilinear sampling.

Same contract as libcpab/pytorch/interpolation.py:12-172 (pytorch layout: data [N,C,W(,H(,D))],
grid [N,ndim,nP] with the first coordinate fastest, result [N,C,*outsize]); forward and backward
are single fused CUDA kernels (cpab_b200_interpolate_forward/backward) instead of 2^ndim gathers
plus autograd.
"""
from __future__ import annotations

import torch

from . import ops


class _InterpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, grid, outsize):
        ctx.save_for_backward(data, grid)
        return ops.interpolate_forward(data, grid, outsize)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        data, grid = ctx.saved_tensors
        need_data, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dgrid, ddata = ops.interpolate_backward(data, grid, grad.contiguous(),
                                                want_dgrid=need_grid, want_ddata=need_data)
        return ddata, dgrid, None


def interpolate(ndim, data, grid, outsize):
    if data.dim() != ndim + 2:
        raise ValueError(f"data must be [n_batch, n_channels, {ndim} spatial dims]")
    if not (data.is_cuda and grid.is_cuda):
        raise RuntimeError("libcpab_b200 runs on CUDA tensors only")
    return _InterpFunction.apply(data, grid, tuple(int(v) for v in outsize))
