"""Tensor-level wrappers over the C ABI: torch owns device memory and streams, nothing else.

Every function takes CUDA tensors, allocates the outputs with torch, enqueues the kernels on
torch's current stream and returns without synchronising.  Shapes follow the reference
(SURVEY.md appendix A.1).  float32 is the production dtype, float64 the check mode.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import CPAB_F32, CPAB_F64, CPAB_FLAG_FAST_GRAD, CPAB_FLAG_FAST_MATH, check, nc_array


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return CPAB_F32
    if t.dtype == torch.float64:
        return CPAB_F64
    raise TypeError(f"libcpab_b200 supports float32/float64 tensors, got {t.dtype}")


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (libcpab_b200 has no CPU path)")
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def n_cells(nc) -> int:
    return int({1: 1, 2: 4, 3: 5}[len(nc)] * int(np.prod(nc)))


def findcellidx(points: torch.Tensor, nc) -> torch.Tensor:
    """points [ndim,nP] -> int32 [nP] (bit-exact with libcpab/core/cpab_ops.cpp:26-190)."""
    points = _req(points, "points")
    ndim, nP = points.shape
    out = torch.empty(nP, dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.load().cpab_b200_findcellidx(_dtype_code(points), ndim, nc_array(nc),
                                                points.data_ptr(), nP, out.data_ptr(), _stream()),
              "findcellidx")
    return out


def theta_to_trels(theta: torch.Tensor, basis_t: torch.Tensor, nc, nsteps: int):
    """theta [n_theta,d], basis_t [d,D] -> (As, Trels), both [n_theta,nC,ndim,ndim+1]."""
    theta = _req(theta, "theta")
    basis_t = _req(basis_t, "basis_t")
    if basis_t.dtype != theta.dtype:
        raise TypeError("theta and basis must have the same dtype")
    ndim = len(nc)
    n_theta, d = theta.shape
    nC = n_cells(nc)
    if tuple(basis_t.shape) != (d, nC * ndim * (ndim + 1)):
        raise ValueError(f"basis_t has shape {tuple(basis_t.shape)}, expected {(d, nC*ndim*(ndim+1))}")
    As = torch.empty((n_theta, nC, ndim, ndim + 1), dtype=theta.dtype, device=theta.device)
    trels = torch.empty_like(As)
    with torch.cuda.device(theta.device):
        check(_lib.load().cpab_b200_theta_to_trels(_dtype_code(theta), ndim, nc_array(nc), int(nsteps),
                                                   n_theta, d, basis_t.data_ptr(), theta.data_ptr(),
                                                   As.data_ptr(), trels.data_ptr(), _stream()),
              "theta_to_trels")
    return As, trels


def expm(A: torch.Tensor) -> torch.Tensor:
    """Batched expm of [n,m,m], m in 2..4 (the reference's pytorch/expm.py as one kernel)."""
    A = _req(A, "A")
    n, m, m2 = A.shape
    if m != m2:
        raise ValueError("expm expects square matrices")
    E = torch.empty_like(A)
    with torch.cuda.device(A.device):
        check(_lib.load().cpab_b200_expm(_dtype_code(A), m, n, A.data_ptr(), E.data_ptr(), _stream()),
              "expm")
    return E


def _points_layout(points: torch.Tensor, n_theta: int):
    """broadcast flag exactly as the reference decides it (pytorch/transformer.cpp:11)."""
    broadcast = int(points.dim() == 3 and points.shape[0] == n_theta)
    if points.dim() not in (2, 3):
        raise ValueError("points must be [ndim,nP] or [n_theta,ndim,nP]")
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[2] if broadcast else points.shape[1]
    if points.dim() == 3 and not broadcast:
        raise ValueError("a 3-D grid must have n_theta as its first dimension")
    return broadcast, int(ndim), int(nP)


def forward(points: torch.Tensor, trels: torch.Tensor, nc, nsteps: int,
            fast_math: bool = False) -> torch.Tensor:
    """cpab_gpu.forward replacement: [n_theta,ndim,nP] transformed points."""
    points = _req(points, "points")
    trels = _req(trels, "trels")
    n_theta = trels.shape[0]
    broadcast, ndim, nP = _points_layout(points, n_theta)
    if ndim != len(nc) or tuple(trels.shape[1:]) != (n_cells(nc), ndim, ndim + 1):
        raise ValueError("trels/points do not match the tessellation")
    if trels.dtype != points.dtype:
        raise TypeError("points and trels must have the same dtype")
    out = torch.empty((n_theta, ndim, nP), dtype=points.dtype, device=points.device)
    flags = CPAB_FLAG_FAST_MATH if fast_math else 0
    with torch.cuda.device(points.device):
        check(_lib.load().cpab_b200_forward(_dtype_code(points), flags, ndim, nc_array(nc), int(nsteps),
                                            n_theta, nP, broadcast, points.data_ptr(),
                                            trels.data_ptr(), out.data_ptr(), _stream()), "forward")
    return out


def backward_jacobian(points: torch.Tensor, As: torch.Tensor, Bs: torch.Tensor, nc,
                      nsteps: int) -> torch.Tensor:
    """cpab_gpu.backward replacement: the reference-layout [d,n_theta,ndim,nP] tensor."""
    points, As, Bs = _req(points, "points"), _req(As, "As"), _req(Bs, "Bs")
    n_theta, d = As.shape[0], Bs.shape[0]
    broadcast, ndim, nP = _points_layout(points, n_theta)
    jac = torch.empty((d, n_theta, ndim, nP), dtype=points.dtype, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.load().cpab_b200_backward_jacobian(_dtype_code(points), ndim, nc_array(nc),
                                                      int(nsteps), n_theta, d, nP, broadcast,
                                                      points.data_ptr(), As.data_ptr(), Bs.data_ptr(),
                                                      jac.data_ptr(), _stream()), "backward_jacobian")
    return jac


def backward_theta(points: torch.Tensor, As: torch.Tensor, basis: torch.Tensor,
                   grad_out: torch.Tensor, nc, nsteps: int, want_dpoints: bool = False,
                   fast_grad: bool = False, redo_count: torch.Tensor = None):
    """Adjoint gradient: dL/dtheta [n_theta,d] (and dL/dpoints [n_theta,ndim,nP] on request).

    Default: every trajectory follows the cell sequence of the reference's float32 RK2 iterates
    (certified, or re-integrated with the reference's arithmetic); `fast_grad=True` skips that.
    `redo_count` (int32 [n_theta], zeroed) receives the number of re-integrated trajectories."""
    points, As = _req(points, "points"), _req(As, "As")
    basis, grad_out = _req(basis, "basis"), _req(grad_out, "grad_out")
    n_theta = As.shape[0]
    D, d = basis.shape
    broadcast, ndim, nP = _points_layout(points, n_theta)
    if tuple(grad_out.shape) != (n_theta, ndim, nP):
        raise ValueError(f"grad_out has shape {tuple(grad_out.shape)}, expected {(n_theta, ndim, nP)}")
    if D != n_cells(nc) * ndim * (ndim + 1):
        raise ValueError("basis does not match the tessellation")
    lib = _lib.load()
    code = _dtype_code(points)
    ws_bytes = lib.cpab_b200_backward_workspace_bytes(code, ndim, nc_array(nc), n_theta, nP)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=points.device)
    dtheta = torch.empty((n_theta, d), dtype=points.dtype, device=points.device)
    dpoints = torch.empty_like(grad_out) if want_dpoints else None
    flags = CPAB_FLAG_FAST_GRAD if fast_grad else 0
    with torch.cuda.device(points.device):
        check(lib.cpab_b200_backward_theta_diag(code, flags, ndim, nc_array(nc), int(nsteps), n_theta, d, nP,
                                                broadcast, points.data_ptr(), As.data_ptr(),
                                                basis.data_ptr(), grad_out.data_ptr(), dtheta.data_ptr(),
                                                dpoints.data_ptr() if want_dpoints else None,
                                                ws.data_ptr(), ws_bytes,
                                                redo_count.data_ptr() if redo_count is not None else None,
                                                _stream()), "backward_theta")
    return dtheta, dpoints


def rk2_cell_trace(points: torch.Tensor, As: torch.Tensor, nc, nsteps: int, mode: int):
    """Cells recorded by the adjoint's first pass (tests / diagnostics): (cells int32
    [n_theta,nsteps,nP], failed uint8 [n_theta,nP]); mode 0 records, 1 certified, 2 reference."""
    points, As = _req(points, "points"), _req(As, "As")
    if points.dtype != torch.float32:
        raise TypeError("rk2_cell_trace is a float32 diagnostic")
    n_theta = As.shape[0]
    broadcast, ndim, nP = _points_layout(points, n_theta)
    lib = _lib.load()
    ws_bytes = lib.cpab_b200_backward_workspace_bytes(CPAB_F32, ndim, nc_array(nc), n_theta, nP)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=points.device)
    cells = torch.empty((n_theta, int(nsteps), nP), dtype=torch.int32, device=points.device)
    failed = torch.zeros((n_theta, nP), dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        check(lib.cpab_b200_rk2_cell_trace(ndim, nc_array(nc), int(nsteps), n_theta, nP, broadcast, int(mode),
                                           points.data_ptr(), As.data_ptr(), ws.data_ptr(), ws_bytes,
                                           cells.data_ptr(), failed.data_ptr(), _stream()), "rk2_cell_trace")
    return cells, failed


def forward_closed_form(points: torch.Tensor, As: torch.Tensor, nc) -> torch.Tensor:
    """Closed-form (hit-time) integration in 1-D / 2-D / 3-D; opt-in extension, not in the reference."""
    points, As = _req(points, "points"), _req(As, "As")
    n_theta = As.shape[0]
    broadcast, ndim, nP = _points_layout(points, n_theta)
    out = torch.empty((n_theta, ndim, nP), dtype=points.dtype, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.load().cpab_b200_forward_closed_form(_dtype_code(points), ndim, nc_array(nc), n_theta, nP,
                                                        broadcast, points.data_ptr(), As.data_ptr(),
                                                        out.data_ptr(), _stream()), "forward_closed_form")
    return out


def closed_form_lane_stats(points: torch.Tensor, As: torch.Tensor, nc):
    """(newpoints, lane utilisation, sub-steps per trajectory) of the 2-D / 3-D hit-time walk: how full
    the warps of the variable-trip-count loop are (tuning key "closed_refill" selects the mitigation)."""
    points, As = _req(points, "points"), _req(As, "As")
    n_theta = As.shape[0]
    broadcast, ndim, nP = _points_layout(points, n_theta)
    out = torch.empty((n_theta, ndim, nP), dtype=points.dtype, device=points.device)
    counts = torch.zeros(2, dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.load().cpab_b200_closed_form_lane_stats(_dtype_code(points), ndim, nc_array(nc), n_theta, nP,
                                                           broadcast, points.data_ptr(), As.data_ptr(),
                                                           out.data_ptr(), counts.data_ptr(), _stream()),
              "closed_form_lane_stats")
    lanes, slots = (int(v) for v in counts.tolist())
    return out, lanes / max(slots, 1), lanes / max(n_theta * nP, 1)


def backward_theta_closed_form(points, As, basis, grad_out, nc, want_dpoints=False, newpoints=None):
    """newpoints: the forward's output, if at hand (2-D / 3-D: spares the kernel its own forward walk)."""
    points, As = _req(points, "points"), _req(As, "As")
    basis, grad_out = _req(basis, "basis"), _req(grad_out, "grad_out")
    if newpoints is not None:
        newpoints = _req(newpoints, "newpoints")
        if newpoints.shape != grad_out.shape or newpoints.dtype != grad_out.dtype:
            raise ValueError("newpoints must have the shape and dtype of grad_out")
    n_theta = As.shape[0]
    D, d = basis.shape
    broadcast, ndim, nP = _points_layout(points, n_theta)
    lib = _lib.load()
    code = _dtype_code(points)
    ws_bytes = lib.cpab_b200_backward_workspace_bytes(code, ndim, nc_array(nc), n_theta, nP)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=points.device)
    dtheta = torch.empty((n_theta, d), dtype=points.dtype, device=points.device)
    dpoints = torch.empty_like(grad_out) if want_dpoints else None
    with torch.cuda.device(points.device):
        check(lib.cpab_b200_backward_theta_closed_form_from(
            code, ndim, nc_array(nc), n_theta, d, nP, broadcast, points.data_ptr(), As.data_ptr(),
            basis.data_ptr(), grad_out.data_ptr(), newpoints.data_ptr() if newpoints is not None else None,
            dtheta.data_ptr(),
            dpoints.data_ptr() if want_dpoints else None, ws.data_ptr(), ws_bytes, _stream()),
            "backward_theta_closed_form")
    return dtheta, dpoints


def interpolate_forward(data: torch.Tensor, grid: torch.Tensor, outsize) -> torch.Tensor:
    """data [N,C,*in], grid [N,ndim,prod(outsize)] -> [N,C,*outsize]."""
    data, grid = _req(data, "data"), _req(grid, "grid")
    ndim = data.dim() - 2
    N, C = data.shape[:2]
    outsize = [int(v) for v in outsize]
    nP = int(np.prod(outsize))
    if ndim not in (1, 2, 3) or len(outsize) != ndim:
        raise ValueError("data must be [N,C,W(,H(,D))] and outsize must have ndim entries")
    if tuple(grid.shape) != (N, ndim, nP):
        raise ValueError(f"grid has shape {tuple(grid.shape)}, expected {(N, ndim, nP)}")
    if grid.dtype != data.dtype:
        raise TypeError("data and grid must have the same dtype")
    out = torch.empty((N, C, *outsize), dtype=data.dtype, device=data.device)
    ins = (ctypes_int_array(data.shape[2:]))
    with torch.cuda.device(data.device):
        check(_lib.load().cpab_b200_interpolate_forward(_dtype_code(data), ndim, N, C, ins,
                                                        ctypes_int_array(outsize), data.data_ptr(),
                                                        grid.data_ptr(), out.data_ptr(), _stream()),
              "interpolate_forward")
    return out


def interpolate_backward(data: torch.Tensor, grid: torch.Tensor, grad_out: torch.Tensor,
                         want_dgrid: bool = True, want_ddata: bool = False):
    data, grid, grad_out = _req(data, "data"), _req(grid, "grid"), _req(grad_out, "grad_out")
    ndim = data.dim() - 2
    N, C = data.shape[:2]
    outsize = [int(v) for v in grad_out.shape[2:]]
    dgrid = torch.empty_like(grid) if want_dgrid else None
    ddata = torch.empty_like(data) if want_ddata else None
    with torch.cuda.device(data.device):
        check(_lib.load().cpab_b200_interpolate_backward(
            _dtype_code(data), ndim, N, C, ctypes_int_array(data.shape[2:]),
            ctypes_int_array(outsize), data.data_ptr(), grid.data_ptr(), grad_out.data_ptr(),
            dgrid.data_ptr() if want_dgrid else None, ddata.data_ptr() if want_ddata else None,
            _stream()), "interpolate_backward")
    return dgrid, ddata


def ctypes_int_array(vals):
    return nc_array([int(v) for v in vals])


def transform_data_forward(points: torch.Tensor, trels: torch.Tensor, data: torch.Tensor, nc,
                           nsteps: int, outsize, fast_math: bool = False):
    """Fused transform_data forward: integrate the shared meshgrid `points` [ndim,nP] for every
    theta and sample `data` [n_theta,C,*in] at the end of each trajectory.  Returns
    (out [n_theta,C,*outsize], grid_t [n_theta,ndim,nP]); identical to forward() + interpolate."""
    points, trels, data = _req(points, "points"), _req(trels, "trels"), _req(data, "data")
    ndim = len(nc)
    n_theta, C = data.shape[:2]
    outsize = [int(v) for v in outsize]
    nP = int(np.prod(outsize))
    if tuple(points.shape) != (ndim, nP):
        raise ValueError(f"points has shape {tuple(points.shape)}, expected {(ndim, nP)}")
    if trels.shape[0] != n_theta or data.dim() != ndim + 2:
        raise ValueError("data, theta batch and tessellation dimension do not match")
    if not (points.dtype == trels.dtype == data.dtype):
        raise TypeError("points, trels and data must have the same dtype")
    out = torch.empty((n_theta, C, *outsize), dtype=data.dtype, device=data.device)
    grid_t = torch.empty((n_theta, ndim, nP), dtype=data.dtype, device=data.device)
    flags = CPAB_FLAG_FAST_MATH if fast_math else 0
    with torch.cuda.device(data.device):
        check(_lib.load().cpab_b200_transform_data_forward(
            _dtype_code(data), flags, ndim, nc_array(nc), int(nsteps), n_theta, C,
            ctypes_int_array(data.shape[2:]), ctypes_int_array(outsize), points.data_ptr(),
            trels.data_ptr(), data.data_ptr(), grid_t.data_ptr(), out.data_ptr(), _stream()),
            "transform_data_forward")
    return out, grid_t


def transform_data_backward(points, As, basis, data, grid_t, grad_out, nc, nsteps: int, fast_grad: bool = False):
    """dL/dtheta of the fused transform_data from the image gradient grad_out [n_theta,C,*outsize]."""
    points, As, basis = _req(points, "points"), _req(As, "As"), _req(basis, "basis")
    data, grid_t, grad_out = _req(data, "data"), _req(grid_t, "grid_t"), _req(grad_out, "grad_out")
    ndim = len(nc)
    n_theta, C = data.shape[:2]
    D, d = basis.shape
    lib = _lib.load()
    code = _dtype_code(data)
    nP = int(grid_t.shape[-1])
    ws_bytes = lib.cpab_b200_backward_workspace_bytes(code, ndim, nc_array(nc), n_theta, nP)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=data.device)
    dtheta = torch.empty((n_theta, d), dtype=data.dtype, device=data.device)
    with torch.cuda.device(data.device):
        check(lib.cpab_b200_transform_data_backward(
            code, CPAB_FLAG_FAST_GRAD if fast_grad else 0, ndim, nc_array(nc), int(nsteps), n_theta, d, C,
            ctypes_int_array(data.shape[2:]),
            ctypes_int_array(grad_out.shape[2:]), points.data_ptr(), As.data_ptr(), basis.data_ptr(),
            data.data_ptr(), grid_t.data_ptr(), grad_out.data_ptr(), dtheta.data_ptr(), ws.data_ptr(),
            ws_bytes, _stream()), "transform_data_backward")
    return dtheta
