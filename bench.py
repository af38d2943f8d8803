#!/usr/bin/env python
"""bench.py -- the measurement contract for libcpab_b200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Metric (BASELINE.json): transformed points x thetas per second ("pairs/s"), forward + backward.

A *step* is one pass of the hot path over one batch of synthetic input:
    out = Cpab.transform_data(data, theta, outsize);  (out * R).sum().backward()  -> theta.grad
i.e. theta -> per-cell affine -> expm -> 50-step integration -> bilinear sampling, and back through
the sampling VJP and the adjoint integration to dL/dtheta.

Workload at N=1 (BASELINE.json configs[1]): 2-D tess_size=[3,3], 64 random thetas ~ N(0,I),
256x256 single-channel images ~ U[0,1), uniform_meshgrid(256,256).  With N ranks every rank
processes its own 64 thetas (theta-sharded, no collective on the path): weak scaling.

One JSON line is printed by rank 0:
  value     pairs/s of the whole job with inputs resident in HBM (device-timed, max over ranks)
  e2e       same metric through the public API with HOST buffers: the pinned->device copy of the
            step's inputs and the device->host read of its result are inside the timed region
            (the upload of step k+1 overlaps the compute of step k on a copy stream)
  roofline  the dominant kernel (adjoint backward): algorithmic FLOPs / measured launch time
            against the FP32 FMA peak measured in this run
  cpu_baseline  the reference's own C++ core (oracle/_ref, else the oracle port) on the host cores
  --impl reference prints the reference arm: the same step computed by the reference's CPU
            implementation on a bounded sample of the workload, all host threads.
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
    "cfg2_2d_t3x3_b64_256x256": ([3, 3], 64, [256, 256], 1, {}),
    "cfg1_1d_t50_b64_1000": ([50], 64, [1000], 1, {}),
    "cfg3_2d_t10x10vp_b512_512x512": ([10, 10], 512, [512, 512], 1, {"volume_perservation": True}),
    "cfg4_3d_t4x4x4_b16_128cubed": ([4, 4, 4], 16, [128, 128, 128], 1, {}),
    "cfg5_1d_t100_b8192_1024": ([100], 8192, [1024], 1, {}),
    # BASELINE configs[4]: CpabSequential of 4 warps, alignment mode -- every series has its own
    # thetas (local gradients) and all series are pulled towards ONE shared, learnable template
    # whose gradient is summed over the theta-shards with an NCCL all-reduce (the only collective)
    "cfg5_1d_seq4_alignment_b8192_1024": ([100], 8192, [1024], 1, {"sequential": 4}),
}
DEFAULT_WORKLOAD = "cfg2_2d_t3x3_b64_256x256"

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
# capture of this very command (profiles/r01_ncu_bench_cfg2_summary.txt); null for workloads
# that were not captured
NCU_TRAFFIC_BYTES = {     # profiles/r01b_ncu_bench_cfg2_summary.txt (fused transform_data kernels)
    "cfg2_2d_t3x3_b64_256x256": {"k_backward": 70.83e6, "k_forward": 19.39e6, "k_interp_fwd": 51.54e6},
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


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
                 "--format=csv,noheader,nounits", "-lms", "100"],
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
def cpu_reference_step(workload: str, cores: int, seed: int = 1236, pts_per_theta: int = 2048):
    """Returns (callable step, pairs per step, description).  One step = forward (expm-stepped)
    + reference-layout Jacobian + contraction for `cores` thetas x `pts_per_theta` grid points
    (a strided sub-sample of the workload's meshgrid), theta-chunked over `cores` threads."""
    from oracle import oracle as O
    from libcpab_b200.tessellation import Tessellation
    tess, n_theta, outsize, C, kw = WORKLOADS[workload]
    ndim = len(tess)
    T = Tessellation(tess, zero_boundary=True, volume_perservation=kw.get("volume_perservation", False))
    rng = np.random.default_rng(seed)
    n_s = max(1, min(n_theta, cores))
    theta = rng.standard_normal((n_s, T.B.shape[1])).astype(np.float32)
    grid = O.uniform_meshgrid(outsize)
    stride = max(1, grid.shape[1] // pts_per_theta)
    pts = np.ascontiguousarray(grid[:, ::stride][:, :pts_per_theta])
    As = O.theta_to_affine(T.B, theta, tess)
    Tr = O.affine_to_trels(As)
    Bs = np.ascontiguousarray(T.B.astype(np.float32).T.reshape(T.B.shape[1], -1, ndim, ndim + 1))
    gout = rng.standard_normal((n_s, ndim, pts.shape[1])).astype(np.float32)
    use_ref = O.have_ref()
    kind = "reference" if use_ref else "port"

    def step():
        if use_ref:
            out = O.ref_forward(pts, Tr, tess, 50, threads=cores)
            jac = O.ref_jacobian(pts, As, Bs, tess, 50, threads=cores)
            g = (jac * gout[None]).sum(axis=(2, 3)).T          # transformer.py:201-202
        else:
            out = O.forward(pts, Tr, tess, 50)
            g = O.theta_grad(pts, As, Bs, gout, tess, 50, threads=cores)
        return out, g

    desc = (f"{n_s} thetas x {pts.shape[1]} points (every {stride}th point of the {outsize} meshgrid), "
            f"forward + theta-Jacobian + contraction, {'libcpab/core/cpab_ops.cpp built as oracle/_ref' if use_ref else 'oracle port'}, "
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
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    sec = time_cpu(step, steps, warmup)
    value = pairs / sec
    tess, n_theta, outsize, C, kw = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "pairs_per_s_fwd_bwd", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "tess_size": tess, "n_theta_per_gpu": n_theta,
                   "outsize": outsize, "channels": C, "nstepsolver": 50, **kw},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcpab_b200 has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from libcpab_b200 import Cpab, _lib, ops

    tess, n_theta, outsize, C, kw = WORKLOADS[args.workload]
    ndim = len(tess)
    nP = int(np.prod(outsize))
    pairs_rank = n_theta * nP
    torch.manual_seed(1234 + 2 + rank)
    kw = dict(kw)
    n_warps = kw.pop("sequential", 0)
    T = Cpab(tess, backend="pytorch", device="gpu", **kw)
    if n_warps:
        run_alignment_arm(args, T, n_warps, tess, n_theta, outsize, C, rank, world, dev)
        return
    theta_h = torch.randn(n_theta, T.params.d).pin_memory()
    data_h = torch.rand(n_theta, C, *outsize).pin_memory()
    R = torch.randn(n_theta, C, *outsize, device=dev)
    theta = theta_h.to(dev).requires_grad_(True)
    data = data_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > L2

    def step_resident():
        theta.grad = None
        out = T.transform_data(data, theta, outsize)
        (out * R).sum().backward()
        return theta.grad

    grad_h = torch.empty(n_theta, T.params.d).pin_memory()
    loss_h = torch.empty(1).pin_memory()

    # End-to-end step: inputs start in PINNED HOST memory every step and the result ends in host
    # memory.  The upload of step k+1 runs on a copy stream while step k computes (two device
    # buffers), the way an input pipeline feeds a training loop; every copy is inside the timed
    # region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [(torch.empty_like(theta_h, device=dev), torch.empty_like(data_h, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # buffer free again
            dev_bufs[slot][0].copy_(theta_h, non_blocking=True)
            dev_bufs[slot][1].copy_(data_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def run_e2e(steps):
        """K pipelined steps; returns device milliseconds for all of them."""
        main = torch.cuda.current_stream()
        for ev in consumed:
            ev.record(main)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        upload(0)
        for k in range(steps):
            slot = k & 1
            if k + 1 < steps:
                upload(slot ^ 1)
            main.wait_event(ready[slot])
            th = dev_bufs[slot][0].detach().requires_grad_(True)
            out = T.transform_data(dev_bufs[slot][1], th, outsize)
            loss = (out * R).sum()
            loss.backward()
            grad_h.copy_(th.grad, non_blocking=True)
            loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
            consumed[slot].record(main)
        b.record(main)
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    def step_e2e():
        run_e2e(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """Device time of `steps` calls (L2 flushed before each, flush not timed); max over ranks."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        barrier()
        for a, b in evs:
            flush.add_(1.0)
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up (also JIT-free: the library is prebuilt; this touches allocator and caches)
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    barrier()

    # FP32 peak of this GPU, measured now
    lib = _lib.load()
    probe_out = torch.zeros(1, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    blocks, iters = 148 * 8, 2048
    probe_ms = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(lib.cpab_b200_fp32_fma_probe(blocks, iters, probe_out.data_ptr(), st), "probe"); b.record()
        torch.cuda.synchronize()
        probe_ms.append(a.elapsed_time(b))
    fp32_peak = blocks * 256 * iters * 64 * 2 / min(probe_ms[1:]) / 1e9       # TFLOP/s

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)
    t0 = time.time()
    total_ms = timed(step_resident, args.steps)
    t1 = time.time()
    launches = _lib.launch_count() - launches0
    prof = {k: _lib.profile_read(k) for k in _lib.PROFILE_SLOTS}
    _lib.profile_enable(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # The HBM-bound kernel of the path, timed on its own (the step above may run the fused
    # transform_data kernels, in which sampling is an epilogue of the integration kernel).
    with torch.no_grad():
        grid0 = T.uniform_meshgrid(outsize)
        grid_t0 = T.transform_grid(grid0, theta.detach())
        ops.interpolate_forward(data, grid_t0, outsize)      # first launch loads the kernel (lazy module loading)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(5):
        flush.add_(1.0)
        ops.interpolate_forward(data, grid_t0, outsize)
    torch.cuda.synchronize()
    interp_prof = _lib.profile_read("interp_fwd")
    _lib.profile_enable(False)
    # ... and once on an HBM-resident problem of the same kind (128 images of 512^2 for 2-D; the
    # workload's own shape is a 20-30 us launch, too short to reach any bandwidth)
    interp_big = None
    if ndim == 2 and rank == 0:
        with torch.no_grad():
            big = [512, 512]
            th2 = torch.randn(128, T.params.d, device=dev)
            gt2 = T.transform_grid(T.uniform_meshgrid(big), th2)
            d2 = torch.rand(128, C, *big, device=dev)
            ops.interpolate_forward(d2, gt2, big)
            torch.cuda.synchronize()
            _lib.profile_enable(True)
            for _ in range(5):
                flush.add_(1.0)
                ops.interpolate_forward(d2, gt2, big)
            torch.cuda.synchronize()
            b_ms, b_n = _lib.profile_read("interp_fwd")
            _lib.profile_enable(False)
            b_bytes = 128 * big[0] * big[1] * (4 * ndim + 8 * C)
            interp_big = {"shape": "128 x %d x 512 x 512" % C, "ms_per_launch": b_ms / max(b_n, 1),
                          "achieved": b_bytes / (b_ms / max(b_n, 1) * 1e-3) / 1e9 if b_n else None, "unit": "GB/s"}
            del th2, gt2, d2

    barrier()
    # two passes of K pipelined steps, the faster one is reported (a pass shares PCIe and the host
    # with whatever else runs on the box; both are measured the same way)
    e2e_local = min(run_e2e(args.steps), run_e2e(args.steps))
    te = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = pairs_rank * world / (ms_per_step * 1e-3)
    e2e_value = pairs_rank * world / (e2e_ms / args.steps * 1e-3)
    peaks, peak_src = measured_peaks()

    bwd_ms, bwd_n = prof["backward"]
    fwd_ms, fwd_n = prof["forward"]
    kshare = {k: (v[0] / total_ms if total_ms else None) for k, v in prof.items()}
    roofline = {
        "kernel": "k_backward (adjoint RK2 sweep)", "bound": "fp32",
        "achieved": pairs_rank * F_BWD[ndim] / (bwd_ms / max(bwd_n, 1) * 1e-3) / 1e12 if bwd_n else None,
        "peak": fp32_peak, "unit": "TFLOP/s",
        "peak_source": "FP32 FMA throughput measured in this run by cpab_b200_fp32_fma_probe "
                       "(MEASURED_PEAKS.json has no FP32 entry; nominal 74.4)",
        "algorithmic_flops_per_pair": F_BWD[ndim], "ms_per_launch": bwd_ms / max(bwd_n, 1) if bwd_n else None,
        "share_of_step": kshare["backward"],
        "traffic": NCU_TRAFFIC_BYTES.get(args.workload, {}).get("k_backward"),
        # fused step: transformed grid + upstream image gradient + 2^n texels + points (compute-bound kernel)
        "algorithmic_bytes": pairs_rank * (4 * ndim + 4 * C + 4 * C) + nP * 4 * ndim,
    }
    roofline["frac"] = roofline["achieved"] / roofline["peak"] if roofline["achieved"] else None
    i_ms, i_n = interp_prof
    interp_bytes = pairs_rank * (4 * ndim + 8 * C)
    roofline_interp = {
        "kernel": "k_interp_fwd", "bound": "hbm",
        "achieved": interp_bytes / (i_ms / max(i_n, 1) * 1e-3) / 1e9 if i_n else None,
        "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "peak_source": peak_src,
        "algorithmic_bytes_per_point": 4 * ndim + 8 * C, "ms_per_launch": i_ms / max(i_n, 1) if i_n else None,
        "share_of_step": kshare["interp_fwd"],
        "traffic": NCU_TRAFFIC_BYTES.get(args.workload, {}).get("k_interp_fwd"),
        "note": "stand-alone k_interp_fwd on this workload's shapes, L2 flushed before each launch "
                "(inside the step the sampler may run fused into the integration kernels)",
    }
    roofline_interp["frac"] = roofline_interp["achieved"] / roofline_interp["peak"] if roofline_interp["achieved"] else None
    if interp_big and interp_big["achieved"]:
        interp_big["frac"] = interp_big["achieved"] / roofline_interp["peak"]
        roofline_interp["hbm_sized"] = interp_big
    roofline_fwd = {
        "kernel": "k_forward", "bound": "fp32",
        "achieved": pairs_rank * F_FWD[ndim] / (fwd_ms / max(fwd_n, 1) * 1e-3) / 1e12 if fwd_n else None,
        "peak": fp32_peak, "unit": "TFLOP/s", "ms_per_launch": fwd_ms / max(fwd_n, 1) if fwd_n else None,
        "share_of_step": kshare["forward"],
    }
    roofline_fwd["frac"] = roofline_fwd["achieved"] / roofline_fwd["peak"] if roofline_fwd["achieved"] else None

    # CPU baseline beside it (rank 0, N=1 only): bounded sample, ~10-30 s of CPU work
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cstep, cpairs, desc, kind = cpu_reference_step(args.workload, cores)
        sec = time_cpu(cstep, 2, 1)
        cpu = {"value": cpairs / sec, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc,
               "seconds_per_sample": sec}

    line = {
        "metric": "pairs_per_s_fwd_bwd", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "tess_size": tess, "n_theta_per_gpu": n_theta,
                   "outsize": outsize, "channels": C, "nstepsolver": 50, "step": "transform_data fwd + bwd wrt theta",
                   "l2": "flushed (256 MiB write) before every timed step (e2e steps are not: their inputs arrive from the host)",
                   "fused_transform_data": bool(T.params.fused_transform_data) if T.params.fused_transform_data is not None
                   else bool(ndim == 1 or n_theta * nP <= (1 << 23)),
                   "parallelism": f"theta-sharded x{world}, no collective", **kw},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_ms / args.steps,
                "how": "public API on pinned host inputs; upload of step k+1 overlaps compute of step k "
                       "(copy stream, 2 device buffers); gradient and loss read back to host every step; "
                       "faster of two passes of K steps",
                "h2d_bytes_per_step": int(theta_h.numel() * 4 + data_h.numel() * 4),
                "d2h_bytes_per_step": int(grad_h.numel() * 4 + 4)},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        "roofline": roofline, "roofline_forward": roofline_fwd, "roofline_interp": roofline_interp,
        "cpu_baseline": cpu, "clocks": clocks,
        "build": lib.cpab_b200_build_info().decode(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_alignment_arm(args, T0, n_warps, tess, n_theta, outsize, C, rank, world, dev):
    """CpabSequential alignment step: forward through n_warps chained flows + interpolation,
    loss against a shared template, backward to every warp's thetas (points_grad extension) and
    to the template, NCCL all-reduce of the template gradient.  pairs = n_warps * n_theta * nP."""
    import torch
    import torch.distributed as dist
    from libcpab_b200 import Cpab, CpabSequential, _lib
    from libcpab_b200.distributed import allreduce_grad_
    Ts = [T0] + [Cpab(tess, backend="pytorch", device="gpu", basis=T0.params.basis) for _ in range(n_warps - 1)]
    for t in Ts:
        t.params.points_grad = True
    S = CpabSequential(*Ts)
    nP = int(np.prod(outsize))
    pairs_rank = n_warps * n_theta * nP
    thetas_h = [(0.5 * torch.randn(n_theta, T0.params.d)).pin_memory() for _ in range(n_warps)]
    data_h = torch.rand(n_theta, C, *outsize).pin_memory()
    thetas = [t.to(dev).requires_grad_(True) for t in thetas_h]
    data = data_h.to(dev)
    template = torch.rand(1, C, *outsize, device=dev).requires_grad_(True)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    grads_h = [torch.empty(n_theta, T0.params.d).pin_memory() for _ in range(n_warps)]

    def step(ths, da):
        for t in ths:
            t.grad = None
        template.grad = None
        out = S.transform_data(da, ths, outsize)
        loss = (out - template).square().sum()
        loss.backward()
        allreduce_grad_(template)
        return loss

    def step_resident():
        step(thetas, data)

    def step_e2e():
        ths = [t.to(dev, non_blocking=True).requires_grad_(True) for t in thetas_h]
        da = data_h.to(dev, non_blocking=True)
        step(ths, da)
        for g, t in zip(grads_h, ths):
            g.copy_(t.grad, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            flush.add_(1.0)
            a.record(); fn(); b.record()
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_resident(); step_e2e()
    l0 = _lib.launch_count()
    _lib.profile_enable(True)
    total_ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    prof = {k: _lib.profile_read(k) for k in _lib.PROFILE_SLOTS}
    _lib.profile_enable(False)
    e2e_ms = timed(step_e2e, args.steps)
    if rank == 0:
        line = {
            "metric": "pairs_per_s_fwd_bwd", "value": pairs_rank * world / (total_ms / args.steps * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "tess_size": tess, "n_theta_per_gpu": n_theta, "outsize": outsize,
                       "warps": n_warps, "step": "CpabSequential.transform_data fwd + bwd to all thetas and a shared template",
                       "collective": "NCCL all-reduce of the shared template gradient (%d floats)" % template.numel(),
                       "l2": "flushed before every timed step", "parallelism": f"theta-sharded x{world}"},
            "e2e": {"value": pairs_rank * world / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(sum(t.numel() for t in thetas_h) * 4 + data_h.numel() * 4),
                    "d2h_bytes_per_step": int(sum(g.numel() for g in grads_h) * 4)},
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
