"""Tessellation geometry and the constraint basis B = null(L) (host precompute, numpy).

Serves the same purpose as libcpab/core/tesselation.py + core/utility.py:27-42 of the reference:
build, once per configuration, the matrix B [D, d] whose columns span the continuous
piecewise-affine velocity fields (optionally zero on the boundary / divergence free).  The
geometry (vertex order, cell numbering) has to agree with the cell search used by the
integrators, so it follows the same conventions:

* 1-D: cell i = [i/nx, (i+1)/nx].
* 2-D: square (ix,iy) -> cells 4*(ix+iy*nx)+{0: top, 1: right, 2: bottom, 3: left}, each the
  triangle (centre, corner, next corner clockwise from the upper-left corner).
* 3-D: cube (ix,iy,iz) -> cells 5*(ix+iy*nx+iz*nx*ny)+{0: central, 1..4: corner tetrahedra};
  cubes of odd parity are rotated a quarter turn about z.

Unlike the reference (an O(nC^2) python loop over cell pairs with set intersections and row-wise
vstack: 126 s for a [4,4,4] tessellation) shared facets are found by sorting integer vertex ids
and the constraint rows are emitted in bulk, so construction is dominated by one dense SVD.

The null-space basis is only defined up to a rotation of its columns; two implementations (or two
LAPACK builds) will not produce the same B.  Parity tests therefore inject the reference's B.
"""
from __future__ import annotations

import hashlib
import itertools
import os

import numpy as np
import scipy.linalg

_CACHE_DIR = os.environ.get("LIBCPAB_B200_BASIS_DIR",
                            os.path.join(os.path.dirname(os.path.abspath(__file__)), "basis_cache"))


def n_cells(nc) -> int:
    return int({1: 1, 2: 4, 3: 5}[len(nc)] * int(np.prod(nc)))


# --------------------------------------------------------------------------------------------
# vertices: [nC, ndim+1, ndim+1] homogeneous coordinates, in the reference's order
# --------------------------------------------------------------------------------------------
def _verts_1d(nc, lo, hi):
    v = np.linspace(lo[0], hi[0], nc[0] + 1)
    out = np.ones((nc[0], 2, 2))
    out[:, 0, 0] = v[:-1]
    out[:, 1, 0] = v[1:]
    return out


def _verts_2d(nc, lo, hi):
    nx, ny = nc
    vx = np.linspace(lo[0], hi[0], nx + 1)
    vy = np.linspace(lo[1], hi[1], ny + 1)
    iy, ix = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")      # ix fastest
    ix, iy = ix.reshape(-1), iy.reshape(-1)
    x0, x1, y0, y1 = vx[ix], vx[ix + 1], vy[iy], vy[iy + 1]
    one = np.ones_like(x0)
    ul, ur = np.stack([x0, y0, one], -1), np.stack([x1, y0, one], -1)
    ll, lr = np.stack([x0, y1, one], -1), np.stack([x1, y1, one], -1)
    ce = np.stack([(x0 + x1) / 2, (y0 + y1) / 2, one], -1)
    tri = np.stack([np.stack([ce, ul, ur], 1), np.stack([ce, ur, lr], 1),
                    np.stack([ce, lr, ll], 1), np.stack([ce, ll, ul], 1)], 1)   # [sq,4,3,3]
    return tri.reshape(-1, 3, 3)


def _verts_3d(nc, lo, hi):
    nx, ny, nz = nc
    vx = np.linspace(lo[0], hi[0], nx + 1)
    vy = np.linspace(lo[1], hi[1], ny + 1)
    vz = np.linspace(lo[2], hi[2], nz + 1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, iy, iz = ix.reshape(-1), iy.reshape(-1), iz.reshape(-1)
    one = np.ones(ix.shape[0])

    def P(dx, dy, dz):
        return np.stack([vx[ix + dx], vy[iy + dy], vz[iz + dz], one], -1)

    ul0, ur0, ll0, lr0 = P(0, 0, 0), P(1, 0, 0), P(0, 1, 0), P(1, 1, 0)
    ul1, ur1, ll1, lr1 = P(0, 0, 1), P(1, 0, 1), P(0, 1, 1), P(1, 1, 1)
    odd = (((ix + iy + iz) % 2) == 1)[:, None]
    # quarter turn about z for odd cubes: (ul,ur,lr,ll) <- (ur,lr,ll,ul)
    ul0, ur0, lr0, ll0 = (np.where(odd, ur0, ul0), np.where(odd, lr0, ur0),
                          np.where(odd, ll0, lr0), np.where(odd, ul0, ll0))
    ul1, ur1, lr1, ll1 = (np.where(odd, ur1, ul1), np.where(odd, lr1, ur1),
                          np.where(odd, ll1, lr1), np.where(odd, ul1, ll1))
    tets = np.stack([np.stack([ll1, ur1, ul0, lr0], 1), np.stack([ul1, ur1, ll1, ul0], 1),
                     np.stack([lr1, ur1, ll1, lr0], 1), np.stack([ll0, ul0, lr0, ll1], 1),
                     np.stack([ur0, ul0, lr0, ur1], 1)], 1)                    # [cube,5,4,4]
    return tets.reshape(-1, 4, 4)


def cell_vertices(nc, domain_min=None, domain_max=None) -> np.ndarray:
    ndim = len(nc)
    lo = [0.0] * ndim if domain_min is None else list(domain_min)
    hi = [1.0] * ndim if domain_max is None else list(domain_max)
    return {1: _verts_1d, 2: _verts_2d, 3: _verts_3d}[ndim](list(nc), lo, hi)


# --------------------------------------------------------------------------------------------
# constraints
# --------------------------------------------------------------------------------------------
def _vertex_ids(verts, nc, lo, hi):
    """Integer id per vertex: vertices live on the half-cell lattice."""
    ndim = len(nc)
    ids = np.zeros(verts.shape[:2], dtype=np.int64)
    for j in range(ndim):
        step = (hi[j] - lo[j]) / nc[j] / 2.0
        q = np.rint((verts[:, :, j] - lo[j]) / step).astype(np.int64)
        ids = ids * (2 * nc[j] + 1) + q
    return ids


def shared_facets(verts, nc, lo, hi):
    """Pairs of cells sharing ndim vertices -> (pairs [m,2], facet vertices [m,ndim,ndim+1])."""
    ndim = len(nc)
    nC = verts.shape[0]
    ids = _vertex_ids(verts, nc, lo, hi)
    combos = list(itertools.combinations(range(ndim + 1), ndim))
    keys, owner, which = [], [], []
    for ci, comb in enumerate(combos):
        k = np.sort(ids[:, comb], axis=1)
        keys.append(k)
        owner.append(np.arange(nC))
        which.append(np.full(nC, ci))
    keys = np.concatenate(keys)
    owner = np.concatenate(owner)
    which = np.concatenate(which)
    order = np.lexsort(keys.T[::-1])
    keys, owner, which = keys[order], owner[order], which[order]
    same = np.all(keys[1:] == keys[:-1], axis=1)
    a, b = np.nonzero(same)[0], np.nonzero(same)[0] + 1
    pairs = np.stack([owner[a], owner[b]], 1)
    combos = np.asarray(combos)
    fverts = verts[owner[a][:, None], combos[which[a]]]
    return pairs, fverts


def _continuity_rows(pairs, fverts, ndim, nC):
    """For every shared vertex v and output row k:  A_i[k,:].v - A_j[k,:].v = 0."""
    ppc = ndim * (ndim + 1)
    m, nv = fverts.shape[:2]
    rows = np.zeros((m, nv, ndim, nC * ppc))
    mi, vi, ki = np.meshgrid(np.arange(m), np.arange(nv), np.arange(ndim), indexing="ij")
    for c in range(ndim + 1):
        col_i = pairs[mi, 0] * ppc + ki * (ndim + 1) + c
        col_j = pairs[mi, 1] * ppc + ki * (ndim + 1) + c
        rows[mi, vi, ki, col_i] = fverts[mi, vi, c]
        rows[mi, vi, ki, col_j] = -fverts[mi, vi, c]
    return rows.reshape(-1, nC * ppc)


def _zero_boundary_rows(verts, lo, hi):
    """Normal velocity component vanishes at every vertex lying on a domain face."""
    nC, nv, m = verts.shape
    ndim = m - 1
    ppc = ndim * m
    if ndim == 1:   # the reference pins only the two end points (tesselation.py:196-200)
        rows = np.zeros((2, nC * ppc))
        rows[0, :2] = [lo[0], 1]
        rows[1, -2:] = [hi[0], 1]
        return rows
    out = []
    for j in range(ndim):
        on = (verts[:, :, j] == lo[j]) | (verts[:, :, j] == hi[j])
        c, v = np.nonzero(on)
        rows = np.zeros((c.shape[0], nC * ppc))
        for col in range(m):
            rows[np.arange(c.shape[0]), c * ppc + j * m + col] = verts[c, v, col]
        out.append(rows)
    return np.concatenate(out)


def _zero_trace_rows(nC, ndim):
    ppc = ndim * (ndim + 1)
    rows = np.zeros((nC, nC * ppc))
    for k in range(ndim):
        rows[np.arange(nC), np.arange(nC) * ppc + k * (ndim + 1) + k] = 1.0
    return rows


def _outside_facets_2d(verts, nc, lo, hi):
    """Auxiliary 'facets' that keep the field continuous outside the domain (valid_outside).

    Same construction as the reference (tesselation.py:240-302): neighbouring boundary triangles
    of the same kind (left/left, right/right, top/top, bottom/bottom) share one vertex; that
    vertex plus a far-away copy of it shifted along the outward axis form a virtual shared edge.
    """
    nx, ny = nc
    pairs, fverts = [], []

    def cell(ix, iy, t):
        return 4 * (ix + iy * nx) + t

    def add(ci, cj, axis):
        vi = {tuple(v) for v in verts[ci]}
        vj = {tuple(v) for v in verts[cj]}
        common = list(vi & vj)
        if len(common) != 1:
            return
        v = np.array(common[0])
        aux = v.copy()
        aux[axis] -= 10
        pairs.append((ci, cj))
        fverts.append(np.stack([v, aux]))

    for iy in range(ny - 1):
        add(cell(0, iy, 3), cell(0, iy + 1, 3), 0)
        add(cell(nx - 1, iy, 1), cell(nx - 1, iy + 1, 1), 0)
    for ix in range(nx - 1):
        add(cell(ix, 0, 0), cell(ix + 1, 0, 0), 1)
        add(cell(ix, ny - 1, 2), cell(ix + 1, ny - 1, 2), 1)
    if not pairs:
        return np.zeros((0, 2), dtype=np.int64), np.zeros((0, 2, 3))
    return np.asarray(pairs), np.asarray(fverts)


def _outside_facets_3d(verts, nc, lo, hi):
    """valid_outside in 3-D (reference: tesselation.py:375-407): two tetrahedra of the SAME cube
    that both have a face on the same domain boundary plane get a virtual shared triangle made of
    their two common vertices and a point pushed one unit out of the domain."""
    nC = verts.shape[0]
    pairs, fverts = [], []
    for cube in range(nC // 5):
        cells = range(5 * cube, 5 * cube + 5)
        for i, j in itertools.combinations(cells, 2):
            for d in range(3):
                low = (np.sum(verts[i][:, d] == lo[d]) == 3) and (np.sum(verts[j][:, d] == lo[d]) == 3)
                high = (np.sum(verts[i][:, d] == hi[d]) == 3) and (np.sum(verts[j][:, d] == hi[d]) == 3)
                if not (low or high):
                    continue
                vi = {tuple(v) for v in verts[i]}
                vj = {tuple(v) for v in verts[j]}
                common = sorted(vi & vj)
                centre = (verts[i][0] + verts[j][0]) / 2.0
                centre[d] += -1 if low else +1
                pts = common + [tuple(centre)]
                if len(pts) != 3:
                    continue
                pairs.append((i, j))
                fverts.append(np.asarray(pts))
    if not pairs:
        return np.zeros((0, 2), dtype=np.int64), np.zeros((0, 3, 4))
    return np.asarray(pairs), np.asarray(fverts)


def constraint_matrix(nc, zero_boundary=True, volume_perservation=False,
                      domain_min=None, domain_max=None) -> np.ndarray:
    ndim = len(nc)
    lo = [0.0] * ndim if domain_min is None else list(domain_min)
    hi = [1.0] * ndim if domain_max is None else list(domain_max)
    verts = cell_vertices(nc, lo, hi)
    nC = verts.shape[0]
    pairs, fverts = shared_facets(verts, nc, lo, hi)
    blocks = [_continuity_rows(pairs, fverts, ndim, nC)]
    if not zero_boundary and ndim >= 2:
        extra = (_outside_facets_2d if ndim == 2 else _outside_facets_3d)(verts, nc, lo, hi)
        if extra[0].shape[0]:
            blocks.append(_continuity_rows(extra[0], extra[1], ndim, nC))
    if zero_boundary:
        blocks.append(_zero_boundary_rows(verts, lo, hi))
    if volume_perservation:
        blocks.append(_zero_trace_rows(nC, ndim))
    return np.concatenate(blocks, axis=0)


def null_space(L: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """Right null space by SVD with the reference's threshold (core/utility.py:27-42)."""
    if L.shape[0] == 0:
        return np.eye(L.shape[1])
    _, s, vh = scipy.linalg.svd(L, full_matrices=True, lapack_driver="gesdd")
    mask = np.concatenate([s <= eps, np.ones(max(0, L.shape[1] - s.shape[0]), dtype=bool)])
    return np.ascontiguousarray(vh[mask].T)


class Tessellation:
    """verts, L, B for one configuration, cached on disk (npz) like the reference's pickle."""

    def __init__(self, nc, domain_min=None, domain_max=None, zero_boundary=True,
                 volume_perservation=False, direc=None, override=False):
        self.nc = [int(v) for v in nc]
        self.ndim = len(self.nc)
        self.domain_min = [0.0] * self.ndim if domain_min is None else list(domain_min)
        self.domain_max = [1.0] * self.ndim if domain_max is None else list(domain_max)
        self.zero_boundary = bool(zero_boundary)
        self.volume_perservation = bool(volume_perservation)
        self.nC = n_cells(self.nc)
        self.n_params = self.ndim * (self.ndim + 1)
        direc = _CACHE_DIR if direc is None else direc
        tag = "cpab_basis_dim%d_tess%s_vo%d_zb%d_vp%d" % (
            self.ndim, "_".join(map(str, self.nc)), int(not self.zero_boundary),
            int(self.zero_boundary), int(self.volume_perservation))
        dom = hashlib.sha1(repr((self.domain_min, self.domain_max)).encode()).hexdigest()[:8]
        self._file = os.path.join(direc, f"{tag}_{dom}.npz")
        self.verts = cell_vertices(self.nc, self.domain_min, self.domain_max)
        if os.path.isfile(self._file) and not override:
            with np.load(self._file) as z:
                self.L, self.B = z["L"], z["B"]
        else:
            self.L = constraint_matrix(self.nc, self.zero_boundary, self.volume_perservation,
                                       self.domain_min, self.domain_max)
            self.B = null_space(self.L)
            try:
                os.makedirs(direc, exist_ok=True)
                tmp = self._file + ".tmp%d.npz" % os.getpid()
                np.savez_compressed(tmp, L=self.L, B=self.B)
                os.replace(tmp, self._file)
            except OSError:
                pass   # read-only install: recompute next time

    def get_cell_centers(self) -> np.ndarray:
        return np.mean(self.verts[:, :, :self.ndim], axis=1)
