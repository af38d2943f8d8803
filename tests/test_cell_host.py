"""The PRODUCT's cell search (libcpab_b200/csrc/cpab_cell.cuh), compiled for the host, swept
against the oracle on the CPU: the exact-arithmetic argument in that header (FMA remainder,
guard-banded diagonal tests, division-free tie-breaks) is checked on adversarial inputs -- cell
faces, diagonals, box faces and corners, and their ulp neighbours -- without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "cell_harness.cpp")
LIB = os.path.join(HERE, "harness", "libcell_harness.so")
HDR = os.path.join(HERE, "..", "libcpab_b200", "csrc", "cpab_cell.cuh")


@pytest.fixture(scope="module")
def harness():
    stale = (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)))
    if stale:
        cpu_fma = "fma" in open("/proc/cpuinfo").read()
        cmd = ["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"]
        if cpu_fma:
            cmd.insert(1, "-mfma")
        subprocess.run(cmd, check=True)
    return ctypes.CDLL(LIB)


def run(lib, pts, nc):
    pts = np.ascontiguousarray(pts)
    out = np.empty(pts.shape[1], np.int32)
    fn = lib.harness_findcellidx_f32 if pts.dtype == np.float32 else lib.harness_findcellidx_f64
    ncarr = np.asarray(nc, np.int32)
    fn(len(nc), ncarr.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p),
       ctypes.c_long(pts.shape[1]), out.ctypes.data_as(ctypes.c_void_p))
    return out


def probes(rng, nc, n, dt):
    ndim = len(nc)
    out = [rng.uniform(-0.2, 1.2, (ndim, n)), rng.uniform(0, 1, (ndim, n))]
    lat = np.stack([rng.integers(0, 4 * nc[j] + 1, n) / (4.0 * nc[j]) for j in range(ndim)])
    latf = lat.astype(dt)
    out += [lat, np.nextafter(latf, dt(2)), np.nextafter(latf, dt(-2))]
    if ndim >= 2:     # on the two diagonals of a square
        cell = [rng.integers(0, nc[j], n) for j in range(ndim)]
        u = rng.uniform(0, 1, n)
        rest = [(cell[2] + rng.uniform(0, 1, n)) / nc[2]] if ndim == 3 else []
        out.append(np.stack([(cell[0] + u) / nc[0], (cell[1] + u) / nc[1]] + rest))
        out.append(np.stack([(cell[0] + u) / nc[0], (cell[1] + 1 - u) / nc[1]] + rest))
    if ndim == 3:     # on the four separating planes of a cube
        u, v = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
        for a, b, c in [(u, v, u + v), (u, v, 2 - u - v), (u, u + v, v), (u + v, u, v)]:
            out.append(np.stack([(cell[0] + a) / nc[0], (cell[1] + b) / nc[1], (cell[2] + np.clip(c, 0, 1)) / nc[2]]))
    face = rng.integers(0, 2, (ndim, n)) + rng.choice([0, 1e-9, -1e-9, 1e-7, -1e-7, 6e-8, 0.05, -0.05], (ndim, n))
    out.append(face)
    mixed = rng.uniform(0, 1, (ndim, n))
    sel = rng.integers(0, ndim, n)
    mixed[sel, np.arange(n)] = rng.integers(0, 2, n) + rng.choice([0, 1e-9, -1e-9, 6e-8, -6e-8], n)
    out.append(mixed)
    return np.ascontiguousarray(np.concatenate(out, axis=1).astype(dt))


@pytest.mark.parametrize("nc", [[1], [2], [7], [50], [100], [1000], [1, 1], [2, 2], [3, 3], [4, 4],
                                [10, 10], [5, 2], [2, 7], [33, 17], [1, 1, 1], [2, 2, 2], [4, 4, 4],
                                [3, 2, 5], [5, 3, 2], [7, 7, 7]])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_product_cell_search_is_bit_exact(harness, nc, dt):
    rng = np.random.default_rng(1000 * len(nc) + nc[0] + (7 if dt is np.float64 else 0))
    pts = probes(rng, nc, 40_000, dt)
    got, ref = run(harness, pts, nc), O.findcellidx(pts, nc)
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, (nc, pts[:, bad[:4]].T.tolist(), got[bad[:4]], ref[bad[:4]])


def test_product_cell_search_on_golden_points(harness):
    from conftest import load_golden
    z = load_golden("cells")
    for key in [k for k in z.files if k.startswith("pts_")]:
        nc = [int(s) for s in key[4:].split("x")]
        assert np.array_equal(run(harness, z[key], nc), z["idx_" + key[4:]]), nc
