"""The hot path under CUDA-graph capture: a whole transform_data forward + backward step is
captured once and replayed -- no host synchronisation, allocation or memset inside the library's
calls (the adjoint kernel's work counters are a self-resetting ring in module memory), so launch-
bound sizes such as BASELINE configs[0] can run as one graph launch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tess,n_theta,size", [([50], 64, [1000]), ([3, 3], 8, [64, 64])])
def test_step_replays_identically_under_a_cuda_graph(tess, n_theta, size):
    from libcpab_b200 import Cpab
    torch.manual_seed(3)
    T = Cpab(tess, backend="pytorch", device="gpu")
    data = torch.rand(n_theta, 1, *size, device="cuda")
    R = torch.randn(n_theta, 1, *size, device="cuda")
    theta = (0.5 * torch.randn(n_theta, T.params.d, device="cuda")).requires_grad_(True)

    def step():
        out = T.transform_data(data, theta, size)
        loss = (out * R).sum()
        (g,) = torch.autograd.grad(loss, theta)
        return out, g

    # everything that touches autograd runs on a side stream: a leaf's grad accumulator remembers
    # the stream it was created on, and the legacy default stream cannot wait on a capturing one
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out_e, g_e = step()                               # eager reference
        for _ in range(2):                                # warm-up (allocator, lazy kernel loading)
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g, g_g = step()
    for _ in range(3):                                    # replays reuse the captured work-counter slot
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_g, out_e)
    # the gradient sums float atomics in a different order from run to run
    assert float((g_g - g_e).abs().max()) <= 2e-5 * float(g_e.abs().max())
    # new inputs through the same graph
    with torch.no_grad():
        theta.mul_(-1.0)
    graph.replay()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        out_e2, g_e2 = step()
    torch.cuda.synchronize()
    assert torch.equal(out_g, out_e2)
    assert float((g_g - g_e2).abs().max()) <= 2e-5 * float(g_e2.abs().max())
