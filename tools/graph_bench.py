#!/usr/bin/env python
"""Eager vs CUDA-graph replay of the transform_data forward + backward step (one B200).
Launch-bound sizes (BASELINE configs[0]) are dominated by ~10 launches and Python glue per step;
captured once, the step is one graph launch.  JSON lines on stdout."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcpab_b200 import Cpab

def run(name, tess, n_theta, size, kw):
    torch.manual_seed(1)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    data = torch.rand(n_theta, 1, *size, device="cuda")
    R = torch.randn(n_theta, 1, *size, device="cuda")
    theta = torch.randn(n_theta, T.params.d, device="cuda").requires_grad_(True)
    def step():
        out = T.transform_data(data, theta, size)
        (g,) = torch.autograd.grad((out * R).sum(), theta)
        return g
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    def timed(fn, n=50):
        with torch.cuda.stream(s):
            for _ in range(5): fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            for _ in range(n): fn()
            b.record(s)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    eager = timed(step)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        step()
    graphed = timed(graph.replay)
    pairs = n_theta * int(np.prod(size))
    print(json.dumps({"workload": name, "eager_ms_per_step": eager, "graph_ms_per_step": graphed,
                      "eager_pairs_per_s": pairs / eager * 1e3, "graph_pairs_per_s": pairs / graphed * 1e3,
                      "note": "back-to-back steps, inputs resident, no L2 flush"}), flush=True)

run("cfg1_1d_t50_b64_1000", [50], 64, [1000], {})
run("cfg2_2d_t3x3_b64_256x256", [3, 3], 64, [256, 256], {})
run("cfg5_1d_t100_b8192_1024", [100], 8192, [1024], {})
