"""Variant sweep of the interpolate kernels (GPU): python tools/interp_variants.py
Prints one JSON line per (shape, direction, variant): ms (median of 10, L2 flushed before each) and GB/s
of the algorithmic bytes (fwd 4n + 8C, bwd-wrt-grid 8n + 8C per point)."""
import json, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from libcpab_b200 import Cpab, ops, _lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, warmup=2, iters=10):
    for _ in range(warmup): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
cases = [([10, 10], 128, [512, 512], {"volume_perservation": True}), ([10, 10], 512, [512, 512], {"volume_perservation": True})]
if "--few" not in sys.argv:
    cases += [([3, 3], 64, [256, 256], {}), ([4, 4, 4], 16, [128, 128, 128], {})]
for tess, n, size, kw in cases:
    T = Cpab(tess, backend='pytorch', device='gpu', **kw)
    torch.manual_seed(1)
    theta = T.sample_transformation(n); grid = T.uniform_meshgrid(size)
    with torch.no_grad(): gt = T.transform_grid(grid, theta)
    data = torch.rand(n, 1, *size, device='cuda'); g2 = torch.randn_like(data)
    nd = len(tess); pts = n * int(np.prod(size))
    for var in (2, 5, 9, 10, 11, 12, 13, 14, 15):
        _lib.set_tuning("interp_variant", var)
        ms = timeit(lambda: ops.interpolate_forward(data, gt, size))
        _lib.profile_enable(True)           # the same launches timed by the library's own event pairs (what bench.py reads)
        for _ in range(5):
            flush.add_(1); ops.interpolate_forward(data, gt, size)
        torch.cuda.synchronize()
        pms, pn = _lib.profile_read("interp_fwd")
        _lib.profile_enable(False)
        print(json.dumps(dict(kind="fwd", tess=tess, n=n, variant=var, ms=round(ms, 4), gbps=round(pts * (4 * nd + 8) / ms / 1e6),
                              lib_event_ms=round(pms / max(pn, 1), 4))), flush=True)
        if var >= 12:
            continue
        ms = timeit(lambda: ops.interpolate_backward(data, gt, g2, True, False))
        print(json.dumps(dict(kind="bwd", tess=tess, n=n, variant=var, ms=round(ms, 4), gbps=round(pts * (8 * nd + 8) / ms / 1e6))), flush=True)
    _lib.set_tuning("interp_variant", 9)
    del data, g2, gt
