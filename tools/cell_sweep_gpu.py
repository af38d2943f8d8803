#!/usr/bin/env python
"""Device cell search against the oracle on a large adversarial sweep (bit-exactness evidence for
the one-sided / packed 2-D search and the packed 3-D search on the GPU itself)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from libcpab_b200 import ops
from oracle import oracle as O
import test_cell_host as H

total = bad_total = 0
t0 = time.time()
for nc in ([3, 3], [10, 10], [25, 25], [7, 4], [100, 100], [1000, 3], [4096, 4096], [1, 5], [33, 17],
           [2, 2, 2], [4, 4, 4], [3, 2, 5], [7, 7, 7], [16, 16, 16], [50], [100], [1000]):
    rng = np.random.default_rng(17 * nc[0] + len(nc))
    for rep in range(4):
        pts = H.probes(rng, nc, 400_000, np.float32)
        # points within a few ulps of every cell face along x, random elsewhere
        k = rng.integers(0, nc[0] + 1, 400_000)
        x = (k.astype(np.float32) * np.float32(1.0 / nc[0]))
        extra = []
        for d in range(-3, 4):
            xx = x.copy()
            for _ in range(abs(d)):
                xx = np.nextafter(xx, np.float32(2 if d > 0 else -2))
            rest = [rng.uniform(0, 1, 400_000).astype(np.float32) for _ in range(len(nc) - 1)]
            extra.append(np.stack([xx] + rest))
        pts = np.ascontiguousarray(np.concatenate([pts] + extra, axis=1))
        got = ops.findcellidx(torch.from_numpy(pts).cuda(), nc).cpu().numpy()
        ref = O.findcellidx(pts, nc)
        nbad = int((got != ref).sum())
        total += pts.shape[1]; bad_total += nbad
    print(json.dumps({"nc": nc, "points_so_far": total, "mismatches_so_far": bad_total}), flush=True)
print(json.dumps({"total_points": total, "mismatches": bad_total, "seconds": time.time() - t0}))
