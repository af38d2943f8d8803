"""Development aid: a numpy transliteration of cpab_closednd.cu's forward walk (float64), to debug its logic on the CPU."""
import sys
import numpy as np
sys.path.insert(0, "tests")
from conftest import load_golden
from oracle import oracle as O
import oracle.oracle as OO

K, NODES, NEWTON = 14, 4, 6
F2 = OO._FACES_2D
F3 = OO._FACES_3D

def horner2(a, t):
    f, df = a[K], 0.0
    for k in range(K - 1, -1, -1):
        df = df * t + f
        f = f * t + a[k]
    return f, df

def walk(x, As, nc, verbose=False):
    ndim = len(nc); spc = 4 if ndim == 2 else 5
    ncv = np.array(nc, float)
    eps = 1e-13 * max(nc)
    x = np.array(x, float)
    idx = np.clip(np.floor(x * ncv), 0, ncv - 1).astype(np.int64)
    typ = OO._simplex_type(ndim, x * ncv - idx, int(idx.sum() & 1) if ndim == 3 else 0)
    trem = 1.0; closed = 0; steps = 0
    while True:
        c = spc * int(idx[0] + idx[1] * nc[0] + (idx[2] * nc[0] * nc[1] if ndim == 3 else 0)) + typ
        A = As[c]
        Lp = A[:, :ndim] * ncv[:, None] / ncv[None, :]
        bp = ncv * (A[:, ndim] + A[:, :ndim] @ (idx / ncv))
        u = x * ncv - idx
        norm = np.abs(Lp).sum(axis=1).max()
        tau = trem
        if norm * tau > 0.5: tau = 0.5 / norm
        ck = np.zeros((K + 1, ndim)); ck[0] = u; ck[1] = Lp @ u + bp
        for k in range(1, K): ck[k + 1] = Lp @ ck[k] / (k + 1)
        parity = int(idx.sum() & 1) if ndim == 3 else 0
        N, D = OO._simplex_faces(ndim, typ, parity)
        best, probe, hit = tau, tau, -1
        for f in range(ndim + 1):
            if (closed >> f) & 1: continue
            n, d = N[f], D[f]
            nzs = np.flatnonzero(n)
            if len(nzs) == 1:
                ax = nzs[0]
                if (n[ax] > 0 and idx[ax] == 0) or (n[ax] < 0 and idx[ax] == nc[ax] - 1): continue
            a = ck @ n; a[0] += d
            a[0] += eps
            lim = best
            r = abs(a[K])
            for k in range(K - 1, 1, -1): r = r * lim + abs(a[k])
            if a[0] + min(0.0, a[1] * lim) - r * lim * lim > 0: continue
            tprev, fprev, dprev = 0.0, a[0], a[1]
            found = a[0] < 0
            cand, after = 0.0, 0.0
            m = 1
            while m <= NODES and not found:
                t = lim if m == NODES else lim * (m / NODES)
                fm, dm = horner2(a, t)
                if fm >= 0 and dprev < 0 and dm > 0:
                    lo, hi = tprev, t
                    for it in range(2 * NEWTON):
                        mid = 0.5 * (lo + hi)
                        fv, dv = horner2(a, mid)
                        if dv < 0: lo = mid
                        else: hi = mid
                    fv, dv = horner2(a, lo)
                    if fv < 0: t, fm, dm = lo, fv, dv
                if fm < 0:
                    found = True
                    lo, hi = tprev, t
                    tt = lo + (hi - lo) * fprev / (fprev - fm)
                    dv = dm
                    for it in range(NEWTON):
                        fv, dv = horner2(a, tt)
                        if fv > 0: lo = tt
                        else: hi = tt
                        tn = tt - fv / dv
                        if not (tn >= lo and tn <= hi): tn = 0.5 * (lo + hi)
                        if tn == tt: break
                        tt = tn
                    cand = tt
                    dt = eps / max(abs(dv), 1e-30)
                    after = cand + dt if cand + dt < t else t
                tprev, fprev, dprev = t, fm, dm
                m += 1
            if found and (cand < best or hit < 0):
                best = min(cand, best); probe = after; hit = f
        at_once = hit >= 0 and not (probe > 0)
        if hit >= 0: best = probe
        tpow = best ** np.arange(K + 1)
        un = tpow @ ck
        x = (idx + un) / ncv
        trem -= best; steps += 1
        if verbose: print("  step", steps, "cell", c, "idx", idx, "typ", typ, "tau", tau, "best", best, "hit", hit, "probe", probe, "x", x)
        done = not (trem > 0)
        if best > 0: closed = 0
        if hit >= 0:
            v = un.copy()
            idx2 = idx.copy(); moved = False
            for j in range(ndim):
                if v[j] < 0 and idx2[j] > 0: idx2[j] -= 1; v[j] += 1; moved = True
                elif v[j] > 1 and idx2[j] < nc[j] - 1: idx2[j] += 1; v[j] -= 1; moved = True
            typ2 = OO._simplex_type(ndim, v, int(idx2.sum() & 1) if ndim == 3 else 0)
            if moved or typ2 != typ:
                idx, typ, closed = idx2, typ2, 0
            elif at_once:
                closed |= 1 << hit
                if verbose: print("    closed", hit)
        if done or steps > 2000: break
    return x

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "d2_t3x3"
    g = load_golden(name); nc = g["nc"].tolist()
    As = O.theta_to_affine(g["B"], g["theta"][-3:].astype(np.float64), nc, np.float64)
    grid = g["grid"].astype(np.float64); grid = grid[:, ::max(1, grid.shape[1] // 150)]
    ref = O.closed_form_nd(grid, As[:1], nc)
    worst = (0, None)
    for i in range(grid.shape[1]):
        got = walk(grid[:, i], As[0], nc)
        e = np.abs(got - ref[0, :, i]).max()
        if e > worst[0]: worst = (e, i)
    print(name, "worst", worst)
    if worst[0] > 1e-10:
        i = worst[1]
        print("point", grid[:, i], "ref", ref[0, :, i])
        walk(grid[:, i], As[0], nc, verbose=True)
