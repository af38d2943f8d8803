#!/usr/bin/env python
"""bench.py -- the measurement contract for libcpab_b200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference_cuda]
                    [--workload NAME] [--no-extras]

Metric (BASELINE.json): transformed points x thetas per second ("pairs/s"), forward + backward.

A *step* is one pass of the hot path over one batch of synthetic input:
    out = Cpab.transform_data(data, theta, outsize);  (out * R).sum().backward()  -> theta.grad
i.e. theta -> per-cell affine -> expm -> 50-step integration -> bilinear sampling, and back through
the sampling VJP and the adjoint integration to dL/dtheta (default gradient mode: certified cell
sequences).

Workload at every N (the largest single-GPU configuration, BASELINE.json configs[2]): 2-D
tess_size=[10,10] volume-preserving, 512 random thetas ~ N(0,I) per GPU, 512x512 single-channel
images ~ U[0,1), uniform_meshgrid(512,512): 134 217 728 pairs, ~3 GB per GPU, larger than L2.  With
N ranks every rank processes its own 512 thetas (theta-sharded, no collective on the path): weak
scaling.

One JSON line is printed by rank 0:
  value      pairs/s of the whole job with inputs resident in HBM (device-timed, max over ranks)
  e2e        same metric through the public API with HOST buffers: the pinned->device copy of the
             step's inputs and the device->host read of its result are inside the timed region
             (the upload of step k+1 overlaps the compute of step k on a copy stream); ONE pass
  roofline   integration + gradient together (k_forward + k_backward + k_backward_redo):
             algorithmic FLOPs / measured launch times against the FP32 FMA peak measured in this
             run; per-kernel breakdown under roofline.kernels
  roofline_interp  the HBM-bound sampling kernels, forward and backward: in the step and on an
             HBM-resident 128 x 512^2 problem, against MEASURED_PEAKS.json:hbm_gbs
  fast_grad  the same step with CPAB_FLAG_FAST_GRAD (no certificate): what the default mode costs
  cpu_baseline     the reference's own C++ core (oracle/_ref, else the oracle port), host cores
  other_workloads  (N=1) compact records of the other BASELINE configs: cfg1, cfg2, cfg4, cfg5
  closed_form      (N=1) the opt-in hit-time integrator (2-D, 3-D): fwd+bwd next to the fixed-step kernels, lane utilisation
  alignment_cfg5   (every N) BASELINE configs[4]: 4-warp CpabSequential alignment step, 8192 series
             x 1024 per GPU, with the NCCL all-reduce of the shared template gradient in the step
  cfg4_point_sharded  (N>1) BASELINE configs[3] with the POINTS split over the ranks and dtheta
             all-reduced: strong scaling of one 16-theta problem
  vs_reference_cuda  (N=1, when oracle/_ref/libcpab_ref_cuda.so exists) the reference's own CUDA
             kernels (libcpab/core/cpab_ops.cu, built unmodified for sm_100a) timed on this GPU
  --impl reference       the reference arm: the same step by the reference's CPU implementation
                         on a bounded sample of the workload, all host threads
  --impl reference_cuda  the reference's CUDA kernels alone (forward on the workload; Jacobian +
                         contraction on a sub-sample whose [d,n_theta,n,nP] tensor fits)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs per (point,theta) pair, SURVEY.md 8-d (FMA = 2, every other fp op = 1):
# forward N(2n(n+1) + C_idx), adjoint backward N(12n^2 + 12n + C_idx), N = 50 steps
F_FWD = {1: 400, 2: 1950, 3: 3550}
F_BWD = {1: 1400, 2: 4950, 3: 9550}

WORKLOADS = {
    # name: (tess, n_theta per GPU, outsize, channels, Cpab kwargs)
    "cfg3_2d_t10x10vp_b512_512x512": ([10, 10], 512, [512, 512], 1, {"volume_perservation": True}),
    "cfg1_1d_t50_b64_1000": ([50], 64, [1000], 1, {}),
    "cfg2_2d_t3x3_b64_256x256": ([3, 3], 64, [256, 256], 1, {}),
    "cfg4_3d_t4x4x4_b16_128cubed": ([4, 4, 4], 16, [128, 128, 128], 1, {}),
    "cfg5_1d_t100_b8192_1024": ([100], 8192, [1024], 1, {}),
    # BASELINE configs[4]: CpabSequential of 4 warps, alignment mode -- every series has its own
    # thetas (local gradients) and all series are pulled towards ONE shared, learnable template
    # whose gradient is summed over the theta-shards with an NCCL all-reduce (the only collective)
    "cfg5_1d_seq4_alignment_b8192_1024": ([100], 8192, [1024], 1, {"sequential": 4}),
}
DEFAULT_WORKLOAD = "cfg3_2d_t10x10vp_b512_512x512"
OTHER_WORKLOADS = ["cfg1_1d_t50_b64_1000", "cfg2_2d_t3x3_b64_256x256", "cfg3_2d_t10x10vp_b512_512x512",
                   "cfg4_3d_t4x4x4_b16_128cubed", "cfg5_1d_t100_b8192_1024"]
ALIGNMENT_WORKLOAD = "cfg5_1d_seq4_alignment_b8192_1024"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full`
    capture of this command (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py);
    {} for workloads that were not captured."""
    try:
        with open(TRAFFIC_FILE) as f:
            return json.load(f).get(workload, {})
    except (OSError, ValueError):
        return {}


# --------------------------------------------------------------------------------------------------
# clocks sampler: nvidia-smi during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or \
               [l.split(", ") for _, l in self.lines]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own core on the host cores, bounded sample of the workload
# --------------------------------------------------------------------------------------------------
def _sample_outsize(outsize, target=2048):
    """A coarser uniform_meshgrid of about `target` points with the aspect of `outsize`."""
    ndim = len(outsize)
    per = max(2, int(round(target ** (1.0 / ndim))))
    if ndim == 1:
        return [min(outsize[0], target)]
    if ndim == 2:
        return [min(outsize[0], 2 * per), min(outsize[1], max(2, target // (2 * per)))]
    return [min(o, per) for o in outsize]


def cpu_reference_step(workload: str, cores: int, seed: int = 1236):
    """Returns (callable step, pairs per step, description, kind).  One step is the SAME step as
    the GPU arm's on a bounded sample -- `cores` thetas, a coarser uniform_meshgrid of ~2048 points,
    images of the workload's own size: theta -> A (basis projection) -> Trels (Pade-13 expm) ->
    forward integration -> linear sampling; backward: sampling VJP -> reference-layout Jacobian
    (libcpab/core/cpab_ops.cpp:262-372) -> contraction (transformer.py:201-202).  The native parts
    run through the reference's own object code (oracle/_ref) when it is present, theta-chunked
    over `cores` threads; the torch-side parts of the reference (projection, expm, interpolation)
    are the oracle's numpy restatements."""
    from oracle import oracle as O
    from libcpab_b200.tessellation import Tessellation
    tess, n_theta, outsize, C, kw = WORKLOADS[workload]
    kw = {k: v for k, v in kw.items() if k != "sequential"}
    ndim = len(tess)
    T = Tessellation(tess, zero_boundary=True, volume_perservation=kw.get("volume_perservation", False))
    rng = np.random.default_rng(seed)
    n_s = max(1, min(n_theta, cores))
    small = _sample_outsize(outsize)
    theta = rng.standard_normal((n_s, T.B.shape[1])).astype(np.float32)
    data = rng.random((n_s, C, *outsize), dtype=np.float32)
    pts = O.uniform_meshgrid(small)
    R = rng.standard_normal((n_s, C, *small)).astype(np.float32)
    Bs = np.ascontiguousarray(T.B.astype(np.float32).T.reshape(T.B.shape[1], -1, ndim, ndim + 1))
    use_ref = O.have_ref()
    kind = "reference" if use_ref else "port"

    def step():
        As = O.theta_to_affine(T.B, theta, tess)                     # transformer.py:146-150
        Tr = O.affine_to_trels(As)                                   # expm.py:11-54
        if use_ref:
            grid_t = O.ref_forward(pts, Tr, tess, 50, threads=cores)
        else:
            grid_t = O.forward(pts, Tr, tess, 50)
        out = O.interpolate(data, grid_t, small)                     # interpolation.py:18-172
        loss = float((out * R).sum())
        dgrid, _ = O.interpolate_vjp(data, grid_t, small, R)
        if use_ref:
            jac = O.ref_jacobian(pts, As, Bs, tess, 50, threads=cores)
            g = (jac * dgrid[None]).sum(axis=(2, 3)).T               # transformer.py:201-202
        else:
            g = O.theta_grad(pts, As, Bs, dgrid, tess, 50, threads=cores)
        return loss, g

    desc = (f"{n_s} thetas x uniform_meshgrid{tuple(small)} = {pts.shape[1]} points sampling {tuple(outsize)} images: "
            f"projection + expm + forward + interpolate + interpolate VJP + theta-Jacobian + contraction; native "
            f"parts = {'libcpab/core/cpab_ops.cpp built as oracle/_ref' if use_ref else 'oracle port'}, "
            f"theta-chunked over {cores} threads")
    return step, n_s * pts.shape[1], desc, kind


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step, pairs, desc, kind = cpu_reference_step(args.workload, cores)
    sec = time_cpu(step, args.steps, args.warmup)
    value = pairs / sec
    tess, n_theta, outsize, C, kw = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "pairs_per_s_fwd_bwd", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "tess_size": tess, "n_theta_per_gpu": n_theta,
                   "outsize": outsize, "channels": C, "nstepsolver": 50,
                   "step": "transform_data fwd + bwd wrt theta", **kw},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU helpers
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide handles of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libcpab_b200 has no CPU path "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # 256 MiB > L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """Device time (ms) of `steps` calls, L2 flushed before each (flush not timed), bracketed by
        barrier + synchronize on both sides; max over ranks."""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in evs:
            self.flush.add_(1.0)
            a.record()
            fn()
            b.record()
        self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))


class Pipeline:
    """End-to-end step runner: inputs start in PINNED HOST memory every step and the result ends in
    host memory.  The upload of step k+1 runs on a copy stream while step k computes (two sets of
    device buffers), the way an input pipeline feeds a training loop; every copy is inside the
    timed region."""

    def __init__(self, ctx, host_inputs, compute, host_outputs):
        torch = ctx.torch
        self.ctx, self.host_inputs, self.compute, self.host_outputs = ctx, host_inputs, compute, host_outputs
        self.copy_stream = torch.cuda.Stream(device=ctx.dev)
        self.bufs = [[torch.empty_like(h, device=ctx.dev) for h in host_inputs] for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.h2d_bytes = int(sum(h.numel() * h.element_size() for h in host_inputs))
        self.d2h_bytes = int(sum(h.numel() * h.element_size() for h in host_outputs))

    def _upload(self, slot):
        torch = self.ctx.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])          # buffer free again
            for d, h in zip(self.bufs[slot], self.host_inputs):
                d.copy_(h, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run(self, steps):
        """K pipelined steps; device milliseconds for all of them (max over ranks)."""
        torch = self.ctx.torch
        main = torch.cuda.current_stream()
        self.ctx.barrier()
        for ev in self.consumed:
            ev.record(main)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        self._upload(0)
        for k in range(steps):
            slot = k & 1
            if k + 1 < steps:
                self._upload(slot ^ 1)
            main.wait_event(self.ready[slot])
            results = self.compute(*self.bufs[slot])
            for h, r in zip(self.host_outputs, results):
                h.copy_(r, non_blocking=True)
            self.consumed[slot].record(main)
        b.record(main)
        self.ctx.barrier()
        return self.ctx.max_over_ranks(a.elapsed_time(b))


def fp32_peak_tflops(ctx):
    """FP32 FMA throughput of this GPU, measured now (MEASURED_PEAKS.json has no FP32 entry)."""
    from libcpab_b200 import _lib
    torch = ctx.torch
    lib = _lib.load()
    out = torch.zeros(1, device=ctx.dev)
    st = torch.cuda.current_stream().cuda_stream
    blocks, iters = 148 * 8, 2048
    ms = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(lib.cpab_b200_fp32_fma_probe(blocks, iters, out.data_ptr(), st), "probe"); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return blocks * 256 * iters * 64 * 2 / min(ms[1:]) / 1e9


def kernel_rooflines(prof, steps, pairs, ndim, C, nP, fp32_peak, hbm_peak, traffic):
    """roofline records from the per-kernel launch times accumulated inside the library."""
    def per_step(slot):
        return prof[slot][0] / steps if prof[slot][1] else 0.0
    t_fwd, t_bwd, t_redo = per_step("forward"), per_step("backward"), per_step("backward_redo")
    t_if, t_ib = per_step("interp_fwd"), per_step("interp_bwd")

    def fp(flops, ms):
        if not ms:
            return None, None
        a = pairs * flops / (ms * 1e-3) / 1e12
        return a, a / fp32_peak
    comb_a, comb_f = fp(F_FWD[ndim] + F_BWD[ndim], t_fwd + t_bwd + t_redo)
    fa, ff = fp(F_FWD[ndim], t_fwd)
    ba, bf = fp(F_BWD[ndim], t_bwd + t_redo)
    roofline = {
        "what": "integration + gradient: k_forward + k_backward (+ k_backward_redo, the certified mode's "
                "reference-arithmetic re-integration) per step",
        "bound": "fp32", "achieved": comb_a, "peak": fp32_peak, "unit": "TFLOP/s", "frac": comb_f,
        "peak_source": "FP32 FMA throughput measured in this run by cpab_b200_fp32_fma_probe "
                       "(MEASURED_PEAKS.json has no FP32 entry; nominal 74.4)",
        "algorithmic_flops_per_pair": F_FWD[ndim] + F_BWD[ndim], "pairs_per_launch": pairs,
        "ms_per_step": t_fwd + t_bwd + t_redo,
        "kernels": {
            "k_forward": {"ms": t_fwd, "flops_per_pair": F_FWD[ndim], "achieved": fa, "frac": ff},
            "k_backward+redo": {"ms": t_bwd + t_redo, "ms_main": t_bwd, "ms_redo": t_redo,
                                "flops_per_pair": F_BWD[ndim], "achieved": ba, "frac": bf},
        },
        "traffic": traffic.get("k_backward"),
        # compute-bound kernels; their HBM traffic for reference: points, upstream gradient, records
        "algorithmic_bytes": pairs * 4 * ndim + nP * 4 * ndim,
    }
    fwd_bytes = pairs * (4 * ndim + 8 * C)                 # grid read + texel read + output write
    bwd_bytes = pairs * (4 * C + 4 * ndim + 4 * C + 4 * ndim)   # grad_out + grid + texels + dgrid
    interp = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s",
              "in_step": {
                  "k_interp_fwd": {"ms": t_if, "bytes_per_point": 4 * ndim + 8 * C,
                                   "achieved": fwd_bytes / (t_if * 1e-3) / 1e9 if t_if else None,
                                   "traffic": traffic.get("k_interp_fwd")},
                  "k_interp_bwd": {"ms": t_ib, "bytes_per_point": 8 * ndim + 8 * C,
                                   "achieved": bwd_bytes / (t_ib * 1e-3) / 1e9 if t_ib else None,
                                   "traffic": traffic.get("k_interp_bwd")}}}
    for k in interp["in_step"].values():
        k["frac"] = k["achieved"] / hbm_peak if k["achieved"] else None
    if not t_if and not t_ib:
        interp["in_step"]["note"] = ("the sampler runs fused inside k_forward / k_backward in this step "
                                     "(transform_data as one forward and one backward kernel)")
    return roofline, interp


def interp_hbm_sized(ctx, T, C):
    """k_interp_fwd / k_interp_bwd alone on an HBM-resident 2-D problem (128 x C x 512 x 512, L2
    flushed before every launch): the bandwidth those kernels reach when nothing else is in the way."""
    from libcpab_b200 import _lib, ops
    torch = ctx.torch
    big = [512, 512]
    with torch.no_grad():
        th = torch.randn(128, T.params.d, device=ctx.dev)
        gt = T.transform_grid(T.uniform_meshgrid(big), th)
        d = torch.rand(128, C, *big, device=ctx.dev)
        g = torch.randn(128, C, *big, device=ctx.dev)
        ops.interpolate_forward(d, gt, big)
        ops.interpolate_backward(d, gt, g, True, False)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(5):
            ctx.flush.add_(1.0)
            ops.interpolate_forward(d, gt, big)
            ctx.flush.add_(1.0)
            ops.interpolate_backward(d, gt, g, True, False)
        torch.cuda.synchronize()
        f_ms, f_n = _lib.profile_read("interp_fwd")
        b_ms, b_n = _lib.profile_read("interp_bwd")
        _lib.profile_enable(False)
    pts = 128 * big[0] * big[1]
    return {"shape": "128 x %d x 512 x 512" % C,
            "k_interp_fwd": {"ms": f_ms / max(f_n, 1), "achieved": pts * (8 + 8 * C) / (f_ms / max(f_n, 1) * 1e-3) / 1e9},
            "k_interp_bwd": {"ms": b_ms / max(b_n, 1), "achieved": pts * (16 + 8 * C) / (b_ms / max(b_n, 1) * 1e-3) / 1e9}}


def run_workload(ctx, name, steps, warmup, fp32_peak, hbm_peak, full=True):
    """One theta-sharded transform_data fwd+bwd workload: device-timed, e2e-timed, kernel split."""
    from libcpab_b200 import Cpab, _lib
    torch = ctx.torch
    tess, n_theta, outsize, C, kw = WORKLOADS[name]
    ndim, nP = len(tess), int(np.prod(outsize))
    pairs_rank = n_theta * nP
    torch.manual_seed(1234 + 2 + ctx.rank)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    theta_h = torch.randn(n_theta, T.params.d).pin_memory()
    data_h = torch.rand(n_theta, C, *outsize).pin_memory()
    R = torch.randn(n_theta, C, *outsize, device=ctx.dev)
    theta = theta_h.to(ctx.dev).requires_grad_(True)
    data = data_h.to(ctx.dev)
    grad_h = torch.empty(n_theta, T.params.d).pin_memory()
    loss_h = torch.empty(1).pin_memory()

    def step_resident():
        theta.grad = None
        out = T.transform_data(data, theta, outsize)
        (out * R).sum().backward()
        return theta.grad

    def compute(th_d, data_d):
        th = th_d.detach().requires_grad_(True)
        out = T.transform_data(data_d, th, outsize)
        loss = (out * R).sum()
        loss.backward()
        return th.grad, loss.detach().reshape(1)

    pipe = Pipeline(ctx, [theta_h, data_h], compute, [grad_h, loss_h])
    for _ in range(max(warmup, 3)):
        step_resident()
    pipe.run(2)
    ctx.barrier()

    sampler = ClockSampler(ctx.local) if (full and ctx.rank == 0) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)
    t0 = time.time()
    total_ms = ctx.timed(step_resident, steps)
    t1 = time.time()
    launches = _lib.launch_count() - launches0
    prof = {k: _lib.profile_read(k) for k in _lib.PROFILE_SLOTS}
    _lib.profile_enable(False)
    clocks = sampler.stop(t0, t1) if sampler else None

    e2e_ms = pipe.run(steps)

    # the same step without the certificate (CPAB_FLAG_FAST_GRAD): what the default mode costs
    T.params.fast_grad = True
    step_resident()
    fast_ms = ctx.timed(step_resident, max(3, steps // 4)) / max(3, steps // 4)
    T.params.fast_grad = False

    ms_per_step = total_ms / steps
    rec = {
        "value": pairs_rank * ctx.world / (ms_per_step * 1e-3), "ms_per_step": ms_per_step,
        "e2e": {"value": pairs_rank * ctx.world / (e2e_ms / steps * 1e-3), "unit": "pairs/s",
                "ms_per_step": e2e_ms / steps, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
        "fast_grad": {"ms_per_step": fast_ms, "value": pairs_rank * ctx.world / (fast_ms * 1e-3),
                      "what": "same step with CPAB_FLAG_FAST_GRAD: no cell-sequence certificate, rare cell flips allowed"},
        "clocks": clocks,
    }
    roofline, interp = kernel_rooflines(prof, steps, pairs_rank, ndim, C, nP, fp32_peak, hbm_peak, ncu_traffic(name))
    rec["roofline"], rec["roofline_interp"] = roofline, interp
    rec["_T"], rec["_shape"] = T, (tess, n_theta, outsize, C, kw)
    del pipe, theta, data, R
    torch.cuda.empty_cache()
    return rec


def run_alignment(ctx, steps, warmup):
    """BASELINE configs[4]: CpabSequential alignment step -- forward through 4 chained flows +
    interpolation, loss against a shared template, backward to every warp's thetas (points_grad
    extension) and to the template, NCCL all-reduce of the template gradient (the collective).
    pairs = n_warps * n_theta * nP per rank.  The all-reduce is bracketed by CUDA events."""
    from libcpab_b200 import Cpab, CpabSequential, _lib
    torch, dist = ctx.torch, ctx.dist
    tess, n_theta, outsize, C, kw = WORKLOADS[ALIGNMENT_WORKLOAD]
    n_warps = kw["sequential"]
    T0 = Cpab(tess, backend="pytorch", device="gpu")
    Ts = [T0] + [Cpab(tess, backend="pytorch", device="gpu", basis=T0.params.basis) for _ in range(n_warps - 1)]
    for t in Ts:
        t.params.points_grad = True
    S = CpabSequential(*Ts)
    nP = int(np.prod(outsize))
    pairs_rank = n_warps * n_theta * nP
    torch.manual_seed(4321 + ctx.rank)
    thetas_h = [(0.5 * torch.randn(n_theta, T0.params.d)).pin_memory() for _ in range(n_warps)]
    data_h = torch.rand(n_theta, C, *outsize).pin_memory()
    thetas = [t.to(ctx.dev).requires_grad_(True) for t in thetas_h]
    data = data_h.to(ctx.dev)
    torch.manual_seed(99)
    template = torch.rand(1, C, *outsize, device=ctx.dev).requires_grad_(True)
    grads_h = [torch.empty(n_theta, T0.params.d).pin_memory() for _ in range(n_warps)]
    tgrad_h = torch.empty(1, C, *outsize).pin_memory()
    ar_events = []

    def step(ths, da):
        template.grad = None
        out = S.transform_data(da, ths, outsize)
        loss = (out - template).square().sum()
        loss.backward()
        if ctx.world > 1:                                   # the only collective of the path
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_reduce(template.grad, op=dist.ReduceOp.SUM)
            b.record()
            ar_events.append((a, b))
        return loss

    def step_resident():
        for t in thetas:
            t.grad = None
        step(thetas, data)

    def compute(*bufs):
        ths = [b.detach().requires_grad_(True) for b in bufs[:n_warps]]
        step(ths, bufs[n_warps])
        return [t.grad for t in ths] + [template.grad]

    pipe = Pipeline(ctx, thetas_h + [data_h], compute, grads_h + [tgrad_h])
    for _ in range(max(warmup, 3)):
        step_resident()
    pipe.run(2)
    ctx.barrier()
    ar_events.clear()
    l0 = _lib.launch_count()
    total_ms = ctx.timed(step_resident, steps)
    launches = _lib.launch_count() - l0
    ar_ms = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ar_events) / max(len(ar_events), 1)) if ar_events else 0.0
    e2e_ms = pipe.run(steps)
    rec = {
        "workload": ALIGNMENT_WORKLOAD, "value": pairs_rank * ctx.world / (total_ms / steps * 1e-3), "unit": "pairs/s",
        "ms_per_step": total_ms / steps, "steps": steps,
        "allreduce_ms": ar_ms, "allreduce_bytes": int(template.numel() * 4),
        "collective": "NCCL all-reduce (sum) of the shared template gradient, on the compute stream after backward"
                      if ctx.world > 1 else "none at N=1 (the all-reduce is skipped for a single rank)",
        "e2e": {"value": pairs_rank * ctx.world / (e2e_ms / steps * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms / steps,
                "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "how": "pinned host thetas + series uploaded on a copy stream one step ahead; all gradients read back"},
        "config": {"tess_size": tess, "warps": n_warps, "series_per_gpu": n_theta, "length": outsize[0],
                   "pairs_per_gpu": pairs_rank, "parallelism": f"theta-sharded x{ctx.world}"},
        "gpu_launches": int(launches),
    }
    del pipe, thetas, data
    torch.cuda.empty_cache()
    return rec


def run_point_sharded(ctx, steps, warmup):
    """BASELINE configs[3] (3-D [4,4,4], 16 thetas, 128^3) with the POINTS of the one problem split
    over the ranks (distributed.PointShardedCpab): every rank integrates its slab of the meshgrid
    for all 16 thetas, samples its slab of the output volume, and dtheta [16, 225] is summed with
    one NCCL all-reduce.  Strong scaling: the total work is fixed."""
    from libcpab_b200 import Cpab
    from libcpab_b200.distributed import PointShardedCpab
    torch, dist = ctx.torch, ctx.dist
    tess, n_theta, outsize, C, kw = WORKLOADS["cfg4_3d_t4x4x4_b16_128cubed"]
    nP = int(np.prod(outsize))
    torch.manual_seed(777)                                  # the SAME problem on every rank
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    P = PointShardedCpab(T)
    theta = torch.randn(n_theta, T.params.d, device=ctx.dev).requires_grad_(True)
    data = torch.rand(n_theta, C, *outsize, device=ctx.dev)
    lo, hi = P.point_bounds(outsize)
    R = torch.randn(n_theta, C, *P.local_outsize(outsize), device=ctx.dev)
    ar_events = []

    def step():
        theta.grad = None
        out = P.transform_data_local(data, theta, outsize)         # [n_theta, C, 128, 128, local slab]
        (out * R).sum().backward()
        if ctx.world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_reduce(theta.grad, op=dist.ReduceOp.SUM)
            b.record()
            ar_events.append((a, b))

    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    ar_events.clear()
    total_ms = ctx.timed(step, steps)
    ar_ms = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ar_events) / max(len(ar_events), 1)) if ar_events else 0.0
    return {"workload": "cfg4_point_sharded", "scaling": "strong", "value": n_theta * nP / (total_ms / steps * 1e-3),
            "unit": "pairs/s", "ms_per_step": total_ms / steps, "steps": steps, "allreduce_ms": ar_ms,
            "allreduce_bytes": int(theta.numel() * 4),
            "config": {"tess_size": tess, "n_theta": n_theta, "outsize": outsize, "points_per_gpu": hi - lo,
                       "parallelism": f"point-sharded x{ctx.world}, all-reduce of dtheta"}}


# --------------------------------------------------------------------------------------------------
# the reference's own CUDA kernels (on-GPU comparator)
# --------------------------------------------------------------------------------------------------
def reference_cuda_records(ctx, name, ours):
    """Times libcpab/core/cpab_ops.cu (built unmodified for sm_100a as oracle/_ref/libcpab_ref_cuda.so
    by oracle/Makefile) with the launch configurations of libcpab/pytorch/transformer_cuda.cu:30-31,
    82-84, on this GPU: forward on the workload, Jacobian + the reference's contraction
    (transformer.py:201) on a sub-sample whose [d, n_theta, ndim, nP] tensor fits in memory."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        return None
    from libcpab_b200 import ops
    from libcpab_b200.transformer import _basis
    torch = ctx.torch
    tess, n_theta, outsize, C, kw = WORKLOADS[name]
    ndim, nP = len(tess), int(np.prod(outsize))
    T = ours["_T"]
    torch.manual_seed(5)
    theta = torch.randn(n_theta, T.params.d, device=ctx.dev)
    grid = T.uniform_meshgrid(outsize)
    B, Bt = _basis(T.params, theta.device, theta.dtype)
    As, Tr = ops.theta_to_trels(theta, Bt, tess, 50)

    def t_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            ctx.flush.add_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    out_ref = torch.empty(n_theta, ndim, nP, device=ctx.dev)
    ref_fwd = t_ms(lambda: ref_cuda.forward(grid, Tr, tess, 50, out_ref))
    our_fwd = t_ms(lambda: ops.forward(grid, Tr, tess, 50))
    # The reference's CUDA build is FMA-contracted and computes fmod as x - floor(x/y)*y
    # (cpab_ops.cu:9-12), so it is NOT bit-identical to the reference's CPU extension (which this
    # library is): a point on or next to a cell face can take the other cell and end elsewhere.
    # Reported as a distribution, not as a pass/fail.
    diff = (ops.forward(grid, Tr, tess, 50) - out_ref).abs().amax(dim=1)
    same = bool(diff.max().item() < 1e-4)
    frac_same = float((diff < 1e-4).float().mean().item())
    max_diff = float(diff.max().item())
    # backward: sub-sample so that d * n_s * ndim * nPs * 4 bytes <= 2 GiB
    d = T.params.d
    n_s = min(n_theta, 16)
    nPs = int(min(nP, (2 << 30) // (4 * d * n_s * ndim)))
    pts = grid[:, :nPs].contiguous()
    gout = torch.randn(n_s, ndim, nPs, device=ctx.dev)
    Bs = B.t().contiguous().view(d, -1, ndim, ndim + 1)
    jac = torch.empty(d, n_s, ndim, nPs, device=ctx.dev)

    def ref_bwd():
        jac.zero_()
        ref_cuda.backward(pts, As[:n_s], Bs, tess, 50, jac)
        return jac.mul_(gout).sum(dim=(2, 3)).t()            # transformer.py:201-202

    ref_bwd_ms = t_ms(ref_bwd)
    our_bwd_ms = t_ms(lambda: ops.backward_theta(pts, As[:n_s].contiguous(), B, gout, tess, 50))
    g_ref = ref_bwd()
    g_our, _ = ops.backward_theta(pts, As[:n_s].contiguous(), B, gout, tess, 50)
    rel = float((g_ref - g_our).abs().max() / g_ref.abs().max())
    return {
        "what": "reference CUDA kernels (libcpab/core/cpab_ops.cu, unmodified, sm_100a) vs this library, same GPU, same inputs",
        "forward": {"pairs": n_theta * nP, "reference_ms": ref_fwd, "ours_ms": our_fwd, "speedup": ref_fwd / our_fwd,
                    "outputs_agree_1e-4": same, "fraction_of_points_within_1e-4": frac_same,
                    "max_abs_diff": max_diff},
        "backward": {"pairs": n_s * nPs, "sample": f"{n_s} thetas x first {nPs} grid points (Jacobian tensor "
                                                     f"{d * n_s * ndim * nPs * 4 / 2**30:.2f} GiB)",
                     "reference_ms": ref_bwd_ms, "ours_ms": our_bwd_ms, "speedup": ref_bwd_ms / our_bwd_ms,
                     "reference_includes": "memset + 3 backward kernels over d + mul_ + sum (transformer.py:187-202)",
                     "gradient_rel_diff": rel},
    }


def closed_form_records(ctx, steps):
    """The opt-in hit-time ("closed-form") integrator of north_star on the 2-D and 3-D BASELINE shapes:
    transform_grid forward + backward w.r.t. theta, device-timed, next to the fixed-step kernels on the same
    inputs, with the lane utilisation of its variable-trip-count loop (refill on / off)."""
    from libcpab_b200 import Cpab, _lib, ops
    torch = ctx.torch
    out = {}
    for name, tess, n_theta, size, kw in (("1d_t100_b8192_1024", [100], 8192, [1024], {}),
                                          ("2d_t10x10vp_b64_512x512", [10, 10], 64, [512, 512], {"volume_perservation": True}),
                                          ("3d_t4x4x4_b16_128cubed", [4, 4, 4], 16, [128, 128, 128], {})):
        torch.manual_seed(77)
        T = Cpab(tess, backend="pytorch", device="gpu", **kw)
        theta = torch.randn(n_theta, T.params.d, device=ctx.dev, requires_grad=True)
        grid = T.uniform_meshgrid(size)
        R = torch.randn(n_theta, len(tess), grid.shape[1], device=ctx.dev)
        pairs = n_theta * grid.shape[1]

        def step():
            theta.grad = None
            (T.transform_grid(grid, theta) * R).sum().backward()
            return theta.grad

        rec = {}
        for mode in ("closed_form", "fixed_step"):
            T.params.closed_form = mode == "closed_form"
            step(); step()
            _lib.profile_enable(True)
            ms = ctx.timed(step, steps) / steps
            prof = {k: _lib.profile_read(k)[0] / steps for k in _lib.PROFILE_SLOTS}
            _lib.profile_enable(False)
            rec[mode] = {"ms_per_step": ms, "value": pairs / (ms * 1e-3), "unit": "pairs/s",
                         "kernel_ms": {k: round(v, 4) for k, v in prof.items() if v}}
        T.params.closed_form = True
        exact = T.transform_grid(grid, theta.detach())
        T.params.closed_form = False
        fixed = T.transform_grid(grid, theta.detach())
        inner = (grid < 1).all(dim=0)
        rec["max_abs_diff_fixed_step_vs_hit_time"] = float((exact - fixed)[:, :, inner].abs().max())
        if len(tess) == 1:          # (the 1-D walk has a closed-form hit time and no lane refill)
            out[name] = rec
            del theta, grid, R
            torch.cuda.empty_cache()
            continue
        B = torch.as_tensor(np.asarray(T.params.basis), dtype=torch.float32, device=ctx.dev)
        As = (B @ theta.detach().T).T.reshape(n_theta, -1, len(tess), len(tess) + 1).contiguous()
        util = {}
        try:
            for refill in (0, 1):
                _lib.set_tuning("closed_refill", refill)
                _, u, per = ops.closed_form_lane_stats(grid, As, tess)
                t = ctx.timed(lambda: ops.forward_closed_form(grid, As, tess), 3) / 3
                util["refill_%d" % refill] = {"lane_utilisation": u, "forward_ms": t}
                rec["substeps_per_trajectory"] = per
        finally:
            _lib.set_tuning("closed_refill", 1)
        rec["divergence"] = util
        out[name] = rec
        del theta, grid, R
        torch.cuda.empty_cache()
    out["what"] = ("opt-in exact integrator (not in the reference): transform_grid fwd + bwd wrt theta, theta ~ N(0, I); "
                   "lane_utilisation = sub-steps executed by lanes / (32 x loop iterations of warps)")
    return out


def run_reference_cuda_arm(args):
    ctx = Ctx()
    if ctx.rank != 0:
        return
    from libcpab_b200 import Cpab
    tess, n_theta, outsize, C, kw = WORKLOADS[args.workload]
    T = Cpab(tess, backend="pytorch", device="gpu", **{k: v for k, v in kw.items() if k != "sequential"})
    rec = reference_cuda_records(ctx, args.workload, {"_T": T})
    if rec is None:
        print(json.dumps({"impl": "reference_cuda", "unavailable": "oracle/_ref/libcpab_ref_cuda.so not built "
                          "(needs /root/reference at build time: make -C oracle)"}), flush=True)
        return
    print(json.dumps({"impl": "reference_cuda", "config": {"workload": args.workload}, **rec}), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    ctx = Ctx()
    from libcpab_b200 import _lib
    lib = _lib.load()
    peaks, peak_src = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    fp32_peak = fp32_peak_tflops(ctx)
    warmup = max(args.warmup, 3)

    if WORKLOADS[args.workload][4].get("sequential"):
        rec = run_alignment(ctx, args.steps, warmup)
        if ctx.rank == 0:
            rec.update({"metric": "pairs_per_s_fwd_bwd", "n_gpus": ctx.world, "warmup": warmup, "higher_is_better": True,
                        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
            rec["config"]["workload"] = args.workload
            print(json.dumps(rec), flush=True)
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return

    main = run_workload(ctx, args.workload, args.steps, warmup, fp32_peak, hbm_peak, full=True)
    T = main.pop("_T")
    tess, n_theta, outsize, C, kw = main.pop("_shape")
    ndim, nP = len(tess), int(np.prod(outsize))

    extras = {}
    if not args.no_extras:
        if ndim == 2 and ctx.rank == 0:
            hs = interp_hbm_sized(ctx, T, C)
            for k in ("k_interp_fwd", "k_interp_bwd"):
                hs[k]["frac"] = hs[k]["achieved"] / hbm_peak
            main["roofline_interp"]["hbm_sized"] = hs
        ctx.barrier()
        # the collective under the driver's eyes, at every N
        extras["alignment_cfg5"] = run_alignment(ctx, max(5, args.steps // 2), warmup)
        if ctx.world > 1:
            extras["cfg4_point_sharded"] = run_point_sharded(ctx, max(5, args.steps // 2), warmup)
        if ctx.world == 1:
            others = {}
            for name in OTHER_WORKLOADS:
                if name == args.workload:
                    continue
                r = run_workload(ctx, name, max(5, args.steps // 2), warmup, fp32_peak, hbm_peak, full=False)
                r.pop("_T"); r.pop("_shape")
                others[name] = {"value": r["value"], "unit": "pairs/s", "ms_per_step": r["ms_per_step"],
                                "e2e_value": r["e2e"]["value"], "fast_grad_ms_per_step": r["fast_grad"]["ms_per_step"],
                                "roofline_frac": r["roofline"]["frac"], "roofline_achieved_tflops": r["roofline"]["achieved"],
                                "kernel_ms_per_step": {k: round(v, 5) for k, v in r["kernel_ms_per_step"].items() if v},
                                "interp_in_step": r["roofline_interp"]["in_step"], "gpu_launches": r["gpu_launches"]}
            extras["other_workloads"] = others
            try:
                extras["closed_form"] = closed_form_records(ctx, max(3, args.steps // 4))
            except Exception as e:
                extras["closed_form"] = {"error": repr(e)}
            try:
                main["_T"] = T
                extras["vs_reference_cuda"] = reference_cuda_records(ctx, args.workload, main)
            except Exception as e:              # the comparator is optional evidence, never the product
                extras["vs_reference_cuda"] = {"error": repr(e)}
            main.pop("_T", None)

    cpu = None
    if ctx.world == 1 and ctx.rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cstep, cpairs, desc, kind = cpu_reference_step(args.workload, cores)
        sec = time_cpu(cstep, 3, 1)
        cpu = {"value": cpairs / sec, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc,
               "seconds_per_sample": sec}

    if ctx.rank == 0:
        fused = bool(T.params.fused_transform_data) if T.params.fused_transform_data is not None else None
        line = {
            "metric": "pairs_per_s_fwd_bwd", "value": main["value"], "unit": "pairs/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "tess_size": tess, "n_theta_per_gpu": n_theta,
                       "outsize": outsize, "channels": C, "nstepsolver": 50, "step": "transform_data fwd + bwd wrt theta",
                       "grad_mode": "default: certified cell sequences (every trajectory follows the reference's cells); "
                                    "fast_grad record = CPAB_FLAG_FAST_GRAD",
                       "l2": "inputs (%.1f GB per GPU) exceed L2; additionally flushed (256 MiB write) before every "
                             "device-timed step" % ((n_theta * nP * 4 * (2 * C + 2 * ndim)) / 1e9),
                       "fused_transform_data": fused,
                       "parallelism": f"theta-sharded x{ctx.world}, no collective", **kw},
            "e2e": dict(main["e2e"], how="public API on pinned host inputs; upload of step k+1 overlaps compute of step k "
                                         "(copy stream, 2 device buffers); gradient and loss read back to host every step; "
                                         "one pass of K steps"),
            "gpu_launches": main["gpu_launches"],
            "kernel_ms_per_step": main["kernel_ms_per_step"],
            "roofline": main["roofline"], "roofline_interp": main["roofline_interp"],
            "fast_grad": main["fast_grad"],
            "cpu_baseline": cpu, "clocks": main["clocks"],
            "build": lib.cpab_b200_build_info().decode(),
        }
        line["roofline_interp"]["peak_source"] = peak_src
        line.update(extras)
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_cuda"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="only the main workload (no alignment / point-sharded / other-workload records)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "reference_cuda":
        run_reference_cuda_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
