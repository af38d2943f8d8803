"""CPU oracle for the CPAB hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package ``libcpab_b200`` never
imports it and has no CPU fallback.

Two native checkers sit behind this module (built by ``oracle/Makefile``):

* ``libcpab_oracle.so``  -- our C99 restatement of the reference's native core
  (``oracle/cpab_oracle.c``), float32 exactly as the reference rounds and an all-double variant.
* ``_ref/libcpab_ref.so`` -- the reference's own ``libcpab/core/cpab_ops.cpp`` compiled where it
  lies under ``/root/reference`` (only the build container has it; the built library travels
  with the repo snapshot).

The pieces of the path that the reference writes in Python/torch are restated here in numpy,
each citing the reference lines it follows (paths relative to ``/root/reference``):

* ``theta_to_affine``   libcpab/pytorch/transformer.py:146-150
* ``expm_pade13``       libcpab/pytorch/expm.py:11-54
* ``uniform_meshgrid``  libcpab/pytorch/functions.py:102-108
* ``interpolate``       libcpab/pytorch/interpolation.py:18-172  (+ its analytic VJP)

Parity status: PINNED by ``tests/test_oracle_pinned.py`` against fixtures generated from the
unmodified reference (``tests/golden/make_golden.py``) and against ``_ref/libcpab_ref.so``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libcpab_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libcpab_ref.so")

_c_int_p = ctypes.POINTER(ctypes.c_int)


def build(quiet: bool = True) -> None:
    """Compile the checkers (``make -C oracle``).  Building the checker is not using it."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _load(path: str) -> ctypes.CDLL:
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


_oracle_lib = None
_ref_lib = None


def oracle_lib() -> ctypes.CDLL:
    global _oracle_lib
    if _oracle_lib is None:
        _oracle_lib = _load(_ORACLE_SO)
    return _oracle_lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib() -> ctypes.CDLL:
    """The reference's own compiled core.  Raises if it was never built (no /root/reference)."""
    global _ref_lib
    if _ref_lib is None:
        if not have_ref():
            build()
        if not have_ref():
            raise FileNotFoundError(_REF_SO + " (reference sources not available to build it)")
        _ref_lib = ctypes.CDLL(_REF_SO)
    return _ref_lib


# --------------------------------------------------------------------------------------------
# geometry helpers shared by everything
# --------------------------------------------------------------------------------------------

def n_cells(nc) -> int:
    """Number of simplices (libcpab/cpab.py:87-97)."""
    ndim = len(nc)
    return int({1: 1, 2: 4, 3: 5}[ndim] * np.prod(nc))


def _suffix(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _nc_arr(nc):
    return np.ascontiguousarray(np.asarray(nc, dtype=np.int32))


# --------------------------------------------------------------------------------------------
# native core: restatement
# --------------------------------------------------------------------------------------------

def findcellidx(points: np.ndarray, nc) -> np.ndarray:
    """points [ndim,nP] -> int32 [nP]; libcpab/core/cpab_ops.cpp:26-190."""
    points = np.ascontiguousarray(points)
    ndim, nP = points.shape
    out = np.empty(nP, dtype=np.int32)
    fn = getattr(oracle_lib(), "cpab_oracle_findcellidx_" + _suffix(points.dtype))
    fn(ctypes.c_int(ndim), _ptr(_nc_arr(nc)), _ptr(points), ctypes.c_long(nP), _ptr(out))
    return out


def forward(points: np.ndarray, trels: np.ndarray, nc, nsteps: int = 50, trace: bool = False):
    """Fixed-step integration; libcpab/core/cpab_ops.cpp:224-260.

    points [ndim,nP] or [n_theta,ndim,nP]; trels [n_theta,nC,ndim,ndim+1] -> [n_theta,ndim,nP]
    (and, with ``trace``, the int32 cell index used at every step, [n_theta,nsteps,nP]).
    """
    trels = np.ascontiguousarray(trels)
    points = np.ascontiguousarray(points, dtype=trels.dtype)
    n_theta = trels.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)  # transformer.cpp:11
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.empty((n_theta, ndim, nP), dtype=trels.dtype)
    cells = np.empty((n_theta, nsteps, nP), dtype=np.int32) if trace else None
    fn = getattr(oracle_lib(), "cpab_oracle_forward_" + _suffix(trels.dtype))
    fn(_ptr(out), _ptr(cells) if trace else None, _ptr(points), _ptr(trels),
       ctypes.c_int(nsteps), _ptr(_nc_arr(nc)), ctypes.c_int(ndim), ctypes.c_long(nP),
       ctypes.c_int(n_theta), ctypes.c_int(broadcast))
    return (out, cells) if trace else out


def jacobian(points: np.ndarray, As: np.ndarray, Bs: np.ndarray, nc, nsteps: int = 50,
             threads: int = 1) -> np.ndarray:
    """theta-Jacobian [d,n_theta,ndim,nP] by the reference's RK2; cpab_ops.cpp:262-372."""
    As = np.ascontiguousarray(As)
    Bs = np.ascontiguousarray(Bs, dtype=As.dtype)
    points = np.ascontiguousarray(points, dtype=As.dtype)
    n_theta, d = As.shape[0], Bs.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    jac = np.zeros((d, n_theta, ndim, nP), dtype=As.dtype)
    fn = getattr(oracle_lib(), "cpab_oracle_jacobian_" + _suffix(As.dtype))
    ncarr = _nc_arr(nc)

    def run(k0, k1):
        fn(_ptr(jac), _ptr(points), _ptr(As), _ptr(Bs), ctypes.c_int(nsteps), _ptr(ncarr),
           ctypes.c_int(n_theta), ctypes.c_int(d), ctypes.c_int(ndim), ctypes.c_long(nP),
           ctypes.c_int(broadcast), ctypes.c_int(0), ctypes.c_int(n_theta),
           ctypes.c_int(k0), ctypes.c_int(k1))

    _split(run, d, threads)
    return jac


def theta_grad(points: np.ndarray, As: np.ndarray, Bs: np.ndarray, gout: np.ndarray, nc,
               nsteps: int = 50, threads: int = 1) -> np.ndarray:
    """dL/dtheta [n_theta,d] = contraction of the RK2 Jacobian with ``gout`` [n_theta,ndim,nP].

    libcpab/core/cpab_ops.cpp:262-372 followed by libcpab/pytorch/transformer.py:201-202,
    without materialising the [d,n_theta,ndim,nP] tensor.  Returned in float64.
    """
    As = np.ascontiguousarray(As)
    Bs = np.ascontiguousarray(Bs, dtype=As.dtype)
    points = np.ascontiguousarray(points, dtype=As.dtype)
    gout = np.ascontiguousarray(gout, dtype=As.dtype)
    n_theta, d = As.shape[0], Bs.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.zeros((n_theta, d), dtype=np.float64)
    fn = getattr(oracle_lib(), "cpab_oracle_theta_grad_" + _suffix(As.dtype))
    ncarr = _nc_arr(nc)

    def run(t0, t1):
        fn(_ptr(out), _ptr(points), _ptr(As), _ptr(Bs), _ptr(gout), ctypes.c_int(nsteps),
           _ptr(ncarr), ctypes.c_int(n_theta), ctypes.c_int(d), ctypes.c_int(ndim),
           ctypes.c_long(nP), ctypes.c_int(broadcast), ctypes.c_int(t0), ctypes.c_int(t1))

    _split(run, n_theta, threads)
    return out


def rk2_flow(points: np.ndarray, As: np.ndarray, nc, nsteps: int = 50) -> np.ndarray:
    """End points of the RK2 trajectories the gradient integrates along; cpab_ops.cpp:289-366."""
    As = np.ascontiguousarray(As)
    points = np.ascontiguousarray(points, dtype=As.dtype)
    n_theta = As.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.empty((n_theta, ndim, nP), dtype=As.dtype)
    fn = getattr(oracle_lib(), "cpab_oracle_rk2_flow_" + _suffix(As.dtype))
    fn(_ptr(out), _ptr(points), _ptr(As), ctypes.c_int(nsteps), _ptr(_nc_arr(nc)),
       ctypes.c_int(n_theta), ctypes.c_int(ndim), ctypes.c_long(nP), ctypes.c_int(broadcast))
    return out


def rk2_trace(points: np.ndarray, As: np.ndarray, nc, nsteps: int = 50):
    """(cells int32 [n_theta,nsteps,nP], end points) of the RK2 trajectories; cpab_ops.cpp:289-366."""
    As = np.ascontiguousarray(As)
    points = np.ascontiguousarray(points, dtype=As.dtype)
    n_theta = As.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.empty((n_theta, ndim, nP), dtype=As.dtype)
    cells = np.empty((n_theta, nsteps, nP), dtype=np.int32)
    fn = getattr(oracle_lib(), "cpab_oracle_rk2_trace_" + _suffix(As.dtype))
    fn(_ptr(cells), _ptr(out), _ptr(points), _ptr(As), ctypes.c_int(nsteps), _ptr(_nc_arr(nc)),
       ctypes.c_int(n_theta), ctypes.c_int(ndim), ctypes.c_long(nP), ctypes.c_int(broadcast))
    return cells, out


def _split(run, n, threads):
    """Run ``run(lo,hi)`` over [0,n) in ``threads`` chunks (ctypes releases the GIL)."""
    threads = max(1, min(int(threads), n))
    if threads == 1:
        run(0, n)
        return
    bounds = np.linspace(0, n, threads + 1).astype(int)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda ab: run(int(ab[0]), int(ab[1])), zip(bounds[:-1], bounds[1:])))


# --------------------------------------------------------------------------------------------
# native core: the reference's own object code (float32 only)
# --------------------------------------------------------------------------------------------

def ref_findcellidx(points: np.ndarray, nc) -> np.ndarray:
    points = np.ascontiguousarray(points, dtype=np.float32)
    ndim, nP = points.shape
    out = np.empty(nP, dtype=np.int32)
    ref_lib().cpab_ref_findcellidx(ctypes.c_int(ndim), _ptr(_nc_arr(nc)), _ptr(points),
                                   ctypes.c_long(nP), _ptr(out))
    return out


def ref_forward(points: np.ndarray, trels: np.ndarray, nc, nsteps: int = 50,
                threads: int = 1) -> np.ndarray:
    """The reference's cpab_forward_op, optionally theta-chunked over host threads."""
    trels = np.ascontiguousarray(trels, dtype=np.float32)
    points = np.ascontiguousarray(points, dtype=np.float32)
    n_theta = trels.shape[0]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.zeros((n_theta, ndim, nP), dtype=np.float32)
    ncarr = _nc_arr(nc)
    lib = ref_lib()

    def run(t0, t1):
        pts = points[t0:t1] if broadcast else points
        lib.cpab_ref_forward(_ptr(out[t0:t1]), _ptr(pts), _ptr(trels[t0:t1]),
                             ctypes.c_int(nsteps), _ptr(ncarr), ctypes.c_int(ndim),
                             ctypes.c_int(nP), ctypes.c_int(t1 - t0), ctypes.c_int(broadcast))

    _split(run, n_theta, threads)
    return out


def ref_jacobian(points: np.ndarray, As: np.ndarray, Bs: np.ndarray, nc, nsteps: int = 50,
                 threads: int = 1) -> np.ndarray:
    """The reference's cpab_backward_op -> [d,n_theta,ndim,nP]; thread-chunked over theta.

    When chunked each chunk writes a private [d,chunk,ndim,nP] block (the reference indexes its
    output with the chunk's own n_theta) which is then copied into place.
    """
    As = np.ascontiguousarray(As, dtype=np.float32)
    Bs = np.ascontiguousarray(Bs, dtype=np.float32)
    points = np.ascontiguousarray(points, dtype=np.float32)
    n_theta, d, nC = As.shape[0], Bs.shape[0], Bs.shape[1]
    broadcast = int(points.ndim == 3 and points.shape[0] == n_theta)
    ndim = points.shape[1] if broadcast else points.shape[0]
    nP = points.shape[-1]
    out = np.zeros((d, n_theta, ndim, nP), dtype=np.float32)
    ncarr = _nc_arr(nc)
    lib = ref_lib()

    def run(t0, t1):
        blk = np.zeros((d, t1 - t0, ndim, nP), dtype=np.float32)
        pts = np.ascontiguousarray(points[t0:t1]) if broadcast else points
        lib.cpab_ref_backward(_ptr(blk), _ptr(pts), _ptr(np.ascontiguousarray(As[t0:t1])),
                              _ptr(Bs), ctypes.c_int(nsteps), _ptr(ncarr),
                              ctypes.c_int(t1 - t0), ctypes.c_int(d), ctypes.c_int(ndim),
                              ctypes.c_int(nP), ctypes.c_int(nC), ctypes.c_int(broadcast))
        out[:, t0:t1] = blk

    _split(run, n_theta, threads)
    return out


# --------------------------------------------------------------------------------------------
# host-side pieces the reference writes in torch, restated in numpy
# --------------------------------------------------------------------------------------------

def theta_to_affine(B: np.ndarray, theta: np.ndarray, nc, dtype=np.float32) -> np.ndarray:
    """As [n_theta,nC,ndim,ndim+1] = (B @ theta.T).T reshaped; transformer.py:146-148.

    The reference converts the float64 basis to float32 first (``torch.Tensor(params.basis)``)
    and multiplies in float32; ``dtype=np.float64`` gives the check-mode variant.
    """
    ndim = len(nc)
    Bc = np.asarray(B).astype(dtype)
    th = np.asarray(theta).astype(dtype)
    A = (Bc @ th.T).T
    return np.ascontiguousarray(A.reshape(th.shape[0], n_cells(nc), ndim, ndim + 1))


_PADE13 = (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
           129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
           40840800., 960960., 16380., 182., 1.)


def expm_pade13(A: np.ndarray) -> np.ndarray:
    """Batched scaling-and-squaring Pade-13 of [n,m,m] in A's own dtype; pytorch/expm.py:11-54."""
    A = np.asarray(A)
    dt = A.dtype
    m = A.shape[-1]
    fro = np.sqrt((np.abs(A) ** 2).sum(axis=(1, 2), keepdims=True)).astype(dt)
    with np.errstate(divide="ignore"):
        lg = (np.log(fro / dt.type(5.371920351148152)) / np.log(dt.type(2.0))).astype(dt)
    nsq = np.maximum(dt.type(0.0), np.ceil(lg)).astype(dt)
    As = (A / (dt.type(2.0) ** nsq)).astype(dt)
    nsq = nsq.reshape(-1).astype(np.int64)
    # the reference builds the coefficients as a float32 tensor and then casts (expm.py:44-48:
    # `torch.Tensor([...]).type(A.dtype)`), so even its float64 expm sees float32-rounded b_k
    b = np.asarray(_PADE13, dtype=np.float32).astype(dt)
    I = np.eye(m, dtype=dt)
    A2 = As @ As
    A4 = A2 @ A2
    A6 = A4 @ A2
    U = As @ (A6 @ (b[13] * A6 + b[11] * A4 + b[9] * A2) + b[7] * A6 + b[5] * A4 + b[3] * A2
              + b[1] * I)
    V = A6 @ (b[12] * A6 + b[10] * A4 + b[8] * A2) + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I
    R = np.linalg.solve((-U + V).astype(dt), (U + V).astype(dt)).astype(dt)
    for i in range(int(nsq.max()) if nsq.size else 0):
        sq = R @ R
        R = np.where((nsq > i)[:, None, None], sq, R)
    return R


def affine_to_trels(As: np.ndarray, nsteps: int = 50) -> np.ndarray:
    """Trels = expm(As/nsteps)[:, :ndim, :]; transformer.py:149-155 (dT*AsSquare in As' dtype)."""
    n_theta, nC, ndim, _ = As.shape
    sq = np.zeros((n_theta * nC, ndim + 1, ndim + 1), dtype=As.dtype)
    sq[:, :ndim, :] = As.reshape(-1, ndim, ndim + 1)
    dT = As.dtype.type(1.0 / nsteps)
    T = expm_pade13((dT * sq).astype(As.dtype))
    return np.ascontiguousarray(T[:, :ndim, :].reshape(n_theta, nC, ndim, ndim + 1))


def uniform_meshgrid(n_points, dtype=np.float32) -> np.ndarray:
    """[ndim,nP] grid over [0,1]^ndim, first coordinate fastest; functions.py:102-108.

    The 1-D ``linspace`` is taken from torch itself (``torch.linspace`` on the CPU, exactly the
    call the reference makes): its vectorised kernel rounds differently from the textbook
    start+i*step in the last bit, and the grid is an INPUT of the path, not part of it.  What is
    restated here is the ordering: meshgrid over the reversed axis list, flattened, reversed back,
    so the first coordinate varies fastest.
    """
    import torch
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    lins = [torch.linspace(0, 1, int(n), dtype=tdt).numpy() for n in n_points]
    mesh = np.meshgrid(*lins[::-1], indexing="ij")
    return np.ascontiguousarray(np.stack([g.reshape(-1) for g in mesh[::-1]], axis=0))


def _taps(g, size):
    """Scale, floor, +1, clamp, weight; interpolation.py:29-47 (and its 2-D/3-D twins)."""
    x = (g * g.dtype.type(size - 1)).astype(g.dtype)
    x0 = np.floor(x).astype(np.int64)
    x1 = x0 + 1
    x0 = np.clip(x0, 0, size - 1)
    x1 = np.clip(x1, 0, size - 1)
    w = (x - x0.astype(g.dtype)).astype(g.dtype)
    return x0, x1, w


def interpolate(data: np.ndarray, grid: np.ndarray, outsize) -> np.ndarray:
    """Linear / bilinear / trilinear sampling; pytorch/interpolation.py:18-172.

    data [N,C,W(,H(,D))], grid [N,ndim,nP] (first coordinate fastest in nP) ->
    [N,C,*outsize] with the reference's reshape/permute (output index order W_o,H_o,D_o).
    Arithmetic is done in data's dtype in the reference's order of operations.
    """
    data = np.asarray(data)
    grid = np.asarray(grid, dtype=data.dtype)
    N, C = data.shape[:2]
    ndim = data.ndim - 2
    one = data.dtype.type(1.0)
    b = np.arange(N)[:, None]
    if ndim == 1:
        x0, x1, xd = _taps(grid[:, 0], data.shape[2])
        xd = xd[..., None]
        c0 = data[b, :, x0]
        c1 = data[b, :, x1]
        c = c0 * (one - xd) + c1 * xd                       # [N,nP,C]
        out = c.reshape(N, outsize[0], C).transpose(0, 2, 1)
    elif ndim == 2:
        x0, x1, xd = _taps(grid[:, 0], data.shape[2])
        y0, y1, yd = _taps(grid[:, 1], data.shape[3])
        xd, yd = xd[..., None], yd[..., None]
        c00 = data[b, :, x0, y0]
        c01 = data[b, :, x0, y1]
        c10 = data[b, :, x1, y0]
        c11 = data[b, :, x1, y1]
        c0 = c00 * (one - xd) + c10 * xd
        c1 = c01 * (one - xd) + c11 * xd
        c = c0 * (one - yd) + c1 * yd
        out = c.reshape(N, outsize[1], outsize[0], C).transpose(0, 3, 2, 1)
    else:
        x0, x1, xd = _taps(grid[:, 0], data.shape[2])
        y0, y1, yd = _taps(grid[:, 1], data.shape[3])
        z0, z1, zd = _taps(grid[:, 2], data.shape[4])
        xd, yd, zd = xd[..., None], yd[..., None], zd[..., None]
        c000 = data[b, :, x0, y0, z0]
        c001 = data[b, :, x0, y0, z1]
        c010 = data[b, :, x0, y1, z0]
        c011 = data[b, :, x0, y1, z1]
        c100 = data[b, :, x1, y0, z0]
        c101 = data[b, :, x1, y0, z1]
        c110 = data[b, :, x1, y1, z0]
        c111 = data[b, :, x1, y1, z1]
        c00 = c000 * (one - xd) + c100 * xd
        c01 = c001 * (one - xd) + c101 * xd
        c10 = c010 * (one - xd) + c110 * xd
        c11 = c011 * (one - xd) + c111 * xd
        c0 = c00 * (one - yd) + c10 * yd
        c1 = c01 * (one - yd) + c11 * yd
        c = c0 * (one - zd) + c1 * zd
        out = c.reshape(N, outsize[2], outsize[1], outsize[0], C).transpose(0, 4, 3, 2, 1)
    return np.ascontiguousarray(out)


def interpolate_vjp(data: np.ndarray, grid: np.ndarray, outsize, gout: np.ndarray):
    """(d/dgrid, d/ddata) of ``sum(interpolate(data,grid)*gout)`` -- what autograd derives from
    pytorch/interpolation.py: floor/clamp are constants, the weight xd = x - x0 carries d/dx =
    (size-1).  Carried in float64 and cast back (the oracle is the better-conditioned side).
    """
    data64 = np.asarray(data, dtype=np.float64)
    grid = np.asarray(grid)
    N, C = data.shape[:2]
    ndim = data.ndim - 2
    nP = grid.shape[-1]
    sizes = data.shape[2:]
    # upstream gradient in point order [N,nP,C]
    perm = (0,) + tuple(range(ndim + 1, 1, -1)) + (1,)
    g = np.asarray(gout, dtype=np.float64).transpose(perm).reshape(N, nP, C)
    taps = [_taps(np.asarray(grid[:, j], dtype=data.dtype), sizes[j]) for j in range(ndim)]
    dgrid = np.zeros((N, ndim, nP), dtype=np.float64)
    ddata = np.zeros_like(data64)
    b = np.broadcast_to(np.arange(N)[:, None], (N, nP))
    for corner in range(1 << ndim):
        bits = [(corner >> j) & 1 for j in range(ndim)]
        idx = tuple(taps[j][bits[j]] for j in range(ndim))
        wts = [taps[j][2].astype(np.float64) if bits[j] else 1.0 - taps[j][2].astype(np.float64)
               for j in range(ndim)]
        w = np.ones((N, nP))
        for j in range(ndim):
            w = w * wts[j]
        val = data64[(b, slice(None)) + idx]                 # [N,nP,C]
        np.add.at(ddata, (b[..., None], np.arange(C)[None, None, :]) +
                  tuple(i[..., None] for i in idx), g * w[..., None])
        gv = (g * val).sum(-1)                               # [N,nP]
        for j in range(ndim):
            wj = np.ones((N, nP))
            for l in range(ndim):
                if l != j:
                    wj = wj * wts[l]
            sign = 1.0 if bits[j] else -1.0
            dgrid[:, j] += sign * wj * gv * (sizes[j] - 1)
    return dgrid.astype(data.dtype), ddata.astype(data.dtype)


# --------------------------------------------------------------------------------------------
# closed-form 1-D integration -- NOT part of the reference
# --------------------------------------------------------------------------------------------

def closed_form_1d(points: np.ndarray, As: np.ndarray, nc) -> np.ndarray:
    """Exact flow of a 1-D CPA field for unit time (hit-time algorithm, Freifeld et al. 2017).

    PARITY UNPINNED: the reference contains no closed-form integrator, so there is
    nothing of the reference to restate or to take golden vectors from.  This float64 numpy
    version is the checker of the opt-in `closed_form` mode; it is anchored to the reference's
    semantics by convergence: the fixed-step scheme (`forward`) tends to it as nstepsolver grows
    (tests/test_closed_form_oracle.py).

    points [1,nP] or [n_theta,1,nP]; As [n_theta,nC,1,2] -> [n_theta,1,nP] (float64).
    """
    As = np.asarray(As, dtype=np.float64)
    n_theta, n = As.shape[0], int(nc[0])
    pts = np.asarray(points, dtype=np.float64)
    x = np.broadcast_to(pts if pts.ndim == 3 else pts[None], (n_theta, 1, pts.shape[-1]))[:, 0].copy()
    a_all, b_all = As[:, :, 0, 0], As[:, :, 0, 1]
    rows = np.arange(n_theta)[:, None]
    t = np.ones_like(x)
    c = np.clip(np.floor(x * n), 0, n - 1).astype(np.int64)
    for _ in range(n + 1):
        a, b = a_all[rows, c], b_all[rows, c]
        v = a * x + b
        right = v > 0
        cn = np.where(right, c + 1, c - 1)
        ok = (v != 0) & (cn >= 0) & (cn < n)
        xb = np.where(right, (c + 1) / n, c / n)
        delta = xb - x
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.where(ok, a * delta / np.where(v == 0, 1.0, v), 0.0)
            L = np.where(np.abs(z) < 1e-8, 1.0 - z / 2, np.log1p(np.where(z > -1, z, 0.0)) / np.where(z == 0, 1.0, z))
            th = np.where(ok & (z > -1), delta / np.where(v == 0, 1.0, v) * L, np.inf)
        cross = th < t
        if not cross.any():
            break
        x = np.where(cross, xb, x)
        t = np.where(cross, t - th, t)
        c = np.where(cross, np.clip(cn, 0, n - 1), c)
    a, b = a_all[rows, c], b_all[rows, c]
    z = a * t
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = np.where(np.abs(z) < 1e-8, 1.0 + z / 2, np.expm1(z) / np.where(z == 0, 1.0, z))
    return (x * np.exp(z) + b * t * phi)[:, None, :]


# --------------------------------------------------------------------------------------------
# closed-form (hit-time) integration in 2-D / 3-D
# --------------------------------------------------------------------------------------------
# Faces of the simplices of one square / cube in its local coordinates u in [0,1]^n, inward positive:
# rows (normal..., offset).  2-D: triangle types of cpab_ops.cpp:94-103; 3-D: tetrahedra of :160-184 in
# the coordinates of an even cube (cubes of odd i+j+k use (x, y) <- (y, 1-x), :170-174).
_FACES_2D = {
    0: [(1, -1, 0), (-1, -1, 1), (0, 1, 0)],
    1: [(1, -1, 0), (1, 1, -1), (-1, 0, 1)],
    2: [(-1, 1, 0), (1, 1, -1), (0, -1, 1)],
    3: [(-1, 1, 0), (-1, -1, 1), (1, 0, 0)],
}
_FACES_3D = {
    0: [(1, 1, -1, 0), (-1, -1, -1, 2), (1, -1, 1, 0), (-1, 1, 1, 0)],
    1: [(-1, -1, 1, 0), (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, -1, 1)],
    2: [(1, 1, 1, -2), (-1, 0, 0, 1), (0, -1, 0, 1), (0, 0, -1, 1)],
    3: [(-1, 1, -1, 0), (1, 0, 0, 0), (0, -1, 0, 1), (0, 0, 1, 0)],
    4: [(1, -1, -1, 0), (-1, 0, 0, 1), (0, 1, 0, 0), (0, 0, 1, 0)],
}


def _simplex_faces(ndim: int, typ: int, parity: int):
    rows = np.array((_FACES_2D if ndim == 2 else _FACES_3D)[typ], dtype=np.float64)
    N, D = rows[:, :ndim].copy(), rows[:, ndim].copy()
    if ndim == 3 and parity:      # a x' + b y' + c z + d with x' = y, y' = 1 - x  =  -b x + a y + c z + (d + b)
        N, D = np.stack([-rows[:, 1], rows[:, 0], rows[:, 2]], axis=1), rows[:, 3] + rows[:, 1]
    return N, D


def _simplex_type(ndim: int, u: np.ndarray, parity: int) -> int:
    """Simplex of a local point (also outside [0,1]^n: the partition continued by its planes)."""
    if ndim == 2:
        x, y = u
        if x < y:
            return 2 if 1 - x < y else 3
        return 1 if 1 - x < y else 0
    x, y, z = u
    if parity:
        x, y = y, 1 - x
    if -x - y + z >= 0:
        return 1
    if x + y + z - 2 >= 0:
        return 2
    if -x + y - z >= 0:
        return 3
    if x - y - z >= 0:
        return 4
    return 0


_GEOMETRY_CHECKED = set()


def _check_geometry(nc) -> None:
    """The face tables against the pinned cell search: on random points of every cube, membership
    by the tables' inequalities is membership by findcellidx (cpab_ops.cpp:33-184)."""
    key = tuple(nc)
    if key in _GEOMETRY_CHECKED:
        return
    ndim = len(nc)
    spc = {2: 4, 3: 5}[ndim]
    rng = np.random.default_rng(11)
    u = rng.uniform(0.01, 0.99, size=(ndim, 4000))
    for cube in range(int(np.prod(nc))):
        idx, s = [], cube
        for j in range(ndim):
            idx.append(s % nc[j]); s //= nc[j]
        parity = (sum(idx) & 1) if ndim == 3 else 0
        pts = (np.array(idx)[:, None] + u) / np.array(nc, dtype=np.float64)[:, None]
        cells = findcellidx(pts, nc)
        for typ in range(spc):
            N, D = _simplex_faces(ndim, typ, parity)
            F = N @ u + D[:, None]
            mine = (F > 1e-9).all(axis=0)
            edge = (np.abs(F) <= 1e-9).any(axis=0)
            theirs = cells == spc * cube + typ
            if ndim == 3:       # sic: the reference clamps z to nz * inc_x (cpab_ops.cpp:141), which cuts the top off
                edge |= pts[2] >= nc[2] / nc[0] - 1e-6      # a tessellation with nz < nx; the tables hold the intended geometry
            if not np.array_equal(mine[~edge], theirs[~edge]):
                raise RuntimeError("simplex face table disagrees with findcellidx (cube %d type %d)" % (cube, typ))
    _GEOMETRY_CHECKED.add(key)


def closed_form_nd(points: np.ndarray, As: np.ndarray, nc, stats: dict | None = None) -> np.ndarray:
    """Exact unit-time flow of a 2-D / 3-D CPA field by the hit-time algorithm of north_star:
    inside a simplex x~(t) = expm(t [[L, b], [0, 0]]) x~; find the time at which the trajectory
    reaches a face, cross into the neighbouring simplex, repeat until t = 1.

    PARITY UNPINNED (as closed_form_1d): the reference has no such integrator.  This float64 checker
    uses scipy's expm and a bracketing root finder (nothing in common with the kernels' Taylor
    polynomials); its simplex geometry is verified against the pinned cell search (_check_geometry)
    and it is anchored by tests/test_closed_form_oracle.py: the float64 RK2 flow of the same field
    (rk2_flow, cpab_ops.cpp:289-366 in double) converges to it at second order.
    Outside the unit box (only trajectories of tessellations without zero boundary get there) the
    field is continued by the planes of the boundary cubes: an outer face is never crossed, the
    other faces are extended -- NOT the reference's rules for outside points (cpab_ops.cpp:47-92,
    119-136), which are discontinuous; `stats["outside"]` marks those trajectories.

    points [n,nP] or [n_theta,n,nP]; As [n_theta,nC,n,n+1] -> [n_theta,n,nP] float64.
    """
    from scipy.linalg import expm
    from scipy.optimize import brentq
    As = np.asarray(As, dtype=np.float64)
    nc = [int(v) for v in nc]
    ndim = len(nc)
    spc = {2: 4, 3: 5}[ndim]
    _check_geometry(nc)
    n_theta = As.shape[0]
    pts = np.asarray(points, dtype=np.float64)
    pts = np.broadcast_to(pts if pts.ndim == 3 else pts[None], (n_theta, ndim, pts.shape[-1]))
    out = np.empty((n_theta, ndim, pts.shape[-1]))
    outside = np.zeros((n_theta, pts.shape[-1]), dtype=bool)
    ncv = np.array(nc, dtype=np.float64)
    eps_on, nodes = 1e-12, 16
    segs = 0
    for th in range(n_theta):
        for i in range(pts.shape[-1]):
            x = pts[th, :, i].copy()
            t_rem = 1.0
            # start simplex: inside the box findcellidx up to ties on faces (_check_geometry), outside it the
            # continuation of the boundary cubes' planes that the walk itself uses
            idx = np.clip(np.floor(x * ncv), 0, ncv - 1).astype(np.int64)
            typ = _simplex_type(ndim, x * ncv - idx, int(idx.sum() & 1) if ndim == 3 else 0)
            closed = set()
            for _ in range(10000):
                parity = int(idx.sum() & 1) if ndim == 3 else 0
                N, D = _simplex_faces(ndim, typ, parity)
                c = spc * int(idx[0] + (idx[1] * nc[0] if ndim > 1 else 0) + (idx[2] * nc[0] * nc[1] if ndim > 2 else 0)) + typ
                At = np.zeros((ndim + 1, ndim + 1))
                At[:ndim] = As[th, c]
                norm = np.abs(At[:ndim, :ndim]).sum(axis=1).max()
                tau = min(t_rem, 0.5 / norm) if norm > 0 else t_rem
                xt = np.append(x, 1.0)

                def local(t):     # the trajectory of this simplex's own flow, local coordinates
                    return (expm(t * At) @ xt)[:ndim] * ncv - idx

                def face(t, p):   # inward-positive face function p and its time derivative
                    e = expm(t * At) @ xt
                    return N[p] @ (e[:ndim] * ncv - idx) + D[p], N[p] @ ((At @ e)[:ndim] * ncv)

                hit_t, hit_p, probe = None, None, None
                for p in range(N.shape[0]):
                    if p in closed:
                        continue
                    axis = np.flatnonzero(N[p])
                    if len(axis) == 1:      # a box face: outer if the cube is at that edge of the domain
                        j = int(axis[0])
                        if (N[p, j] > 0 and idx[j] == 0) or (N[p, j] < 0 and idx[j] == nc[j] - 1):
                            continue
                    lim = tau if hit_t is None else hit_t
                    ts = np.linspace(0.0, lim, nodes + 1)
                    fd = [face(t, p) for t in ts]
                    # exit = the first time f(t) + eps turns negative (a point on the face is still inside)
                    if fd[0][0] < -eps_on:          # already outside: leave at once
                        if hit_t is None or 0.0 < hit_t:
                            hit_t, hit_p, probe = 0.0, p, 0.0
                        continue
                    for m in range(1, nodes + 1):
                        t_m, (f_m, d_m) = ts[m], fd[m]
                        if f_m >= -eps_on and fd[m - 1][1] < 0 and d_m > 0:      # a minimum in between
                            t_min = brentq(lambda t: face(t, p)[1], ts[m - 1], ts[m], xtol=1e-15, rtol=1e-15)
                            if face(t_min, p)[0] < -eps_on:
                                t_m, (f_m, d_m) = t_min, face(t_min, p)
                        if f_m < -eps_on:
                            t_p = brentq(lambda t: face(t, p)[0] + eps_on, ts[m - 1], t_m, xtol=1e-15, rtol=1e-15)
                            after = min(t_m, t_p + eps_on / max(abs(face(t_p, p)[1]), 1e-300))
                            if hit_t is None or t_p < hit_t:
                                hit_t, hit_p, probe = t_p, p, after
                            break
                # on a crossing go to the probe time, a few eps behind the face: strictly inside the simplex the point
                # is classified into below (also when that is this one again: a graze)
                step = tau if hit_t is None else probe
                x_new = (expm(step * At) @ xt)[:ndim]
                t_rem -= step
                segs += 1
                if step > 0:
                    closed = set()
                if (x_new < -1e-9).any() or (x_new > 1 + 1e-9).any():
                    outside[th, i] = True
                if hit_t is not None:
                    u = local(probe)
                    idx2 = idx.copy()
                    for j in range(ndim):
                        if u[j] < 0 and idx2[j] > 0:
                            idx2[j] -= 1; u[j] += 1
                        elif u[j] > 1 and idx2[j] < nc[j] - 1:
                            idx2[j] += 1; u[j] -= 1
                    typ2 = _simplex_type(ndim, u, int(idx2.sum() & 1) if ndim == 3 else 0)
                    if typ2 != typ or (idx2 != idx).any():
                        idx, typ, closed = idx2, typ2, set()
                    elif not probe > 0:
                        closed.add(hit_p)      # outside a face at time 0 yet classified into this simplex: rounding
                x = x_new
                if t_rem <= 0:
                    break
            out[th, :, i] = x
    if stats is not None:
        stats["segments"] = segs
        stats["outside"] = outside
    return out
