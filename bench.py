#!/usr/bin/env python
"""bench.py -- DRRs/sec fwd+bwd (512^3 CT -> 256^2 detector) and achieved algorithmic HBM GB/s vs the B200 peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic poses: Siddon forward (line integrals of
`--batch` poses through a 512^3 fp32 volume onto a 256^2 detector) + Siddon backward (gradients w.r.t. the ray
end points / ray lengths, i.e. the pose-gradient path of 2D/3D registration).  Since every pixel depends on one ray
only, the forward walk also accumulates the ray's 6 end-point sensitivities (they do not depend on the upstream
gradient) and the backward pass is an elementwise kernel: one walk per step instead of two, same outputs.
Reported on ONE JSON line:

  value      whole-job DRRs/s with the rays already resident in HBM (kernels only, CUDA-event timed)
  e2e        the same metric through the public module call `DRR(rot, xyz)` + backward, with the pose parameters
             coming from pinned HOST memory every step and the images / loss / pose gradients copied back
  roofline   algorithmic bytes of the dominant kernel / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (C port of the reference algorithm, OpenMP) on a bounded sample, rank 0 only

With N > 1 (torchrun, one rank per GPU, NCCL) the poses are sharded across ranks (weak scaling: `--batch` poses per
GPU), the volume is replicated, and every step ends with ONE all_gather of the image stack.
`--impl reference` times the CPU oracle port instead (the reference is pure Python and cannot travel to the GPU
box) with all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

VOL = 512
DET = 256
METRIC = "DRRs/sec fwd+bwd (512^3 CT -> 256^2 det)"
WORKLOAD = "siddon fwd+bwd(pose), 512^3 fp32 CT -> 256^2 detector"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the default workload
# (16 poses, 512^3 -> 256^2): profiles/r01_fwd_slab_B16_ncu_summary.txt and gpurun capture of the backward kernel.
# Only quoted when the run uses that workload; otherwise null.
NCU_TRAFFIC_BYTES = {"siddon_fwd_slab_kernel": 1.05e9 + 0.02e9, "siddon_bwd_slab_kernel": 2.26e9 + 0.09e9,
                     "siddon_sens_slab_kernel": 1.38e9 + 0.14e9}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="poses per GPU per step")
    ap.add_argument("--vol", type=int, default=VOL)
    ap.add_argument("--det", type=int, default=DET)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="report the eager end-to-end step only")
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons in the background during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thr = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) == 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def make_host_rays(vol_n, det_n, batch, seed):
    """Rays of `batch` synthetic poses in voxel coordinates, computed with the host-side geometry (CPU torch)."""
    from diffdrr_b200 import DRR, synthetic
    from diffdrr_b200.pose import convert

    subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
    subj.volume.affine = synthetic.make_affine(vol_n)
    drr = DRR(subj, **synthetic.detector_kwargs(det_n))
    rot, xyz = synthetic.make_poses(batch, seed=seed)
    with torch.no_grad():
        src, tgt = drr.detector(convert(rot, xyz, parameterization="euler_angles", convention="ZXY"), None)
        raylen = (tgt - src).norm(dim=-1).unsqueeze(1)
        src, tgt = drr.affine_inverse(src), drr.affine_inverse(tgt)
    return src.numpy(), tgt.numpy(), raylen.numpy()


def cpu_oracle_time(vol_n, det_n, n_drr, repeats, threads):
    """Seconds per fwd+bwd DRR of the C oracle (reference algorithm: all planes merged/sorted per ray, fp32)."""
    from diffdrr_b200 import synthetic
    from oracle import oracle

    oracle.set_threads(threads)
    vol = synthetic.make_volume(vol_n, "rand", seed=0)
    src, tgt, raylen = make_host_rays(vol_n, det_n, max(n_drr, 2), seed=0)
    src, tgt, raylen = src[:n_drr], tgt[:n_drr], raylen[:n_drr]
    gout = np.ones_like(raylen)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.siddon_fwd(vol, src, tgt, raylen, dtype=np.float32)
        oracle.siddon_bwd(vol, src, tgt, raylen, gout, want_vol=False, dtype=np.float32)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank, world):
    """`--impl reference`: the CPU port of the reference's algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle

    threads = oracle.max_threads()
    n_drr = 1
    times = cpu_oracle_time(args.vol, args.det, n_drr, args.warmup + args.steps, threads)[args.warmup:]
    sec = float(np.sum(times))
    value = n_drr * len(times) / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "DRRs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "volume": [args.vol] * 3, "detector": [args.det] * 2,
                   "sample": f"{n_drr} DRR (pose) fwd+bwd per step", "renderer": "siddon"},
        "cpu_baseline": {"value": value, "unit": "DRRs/s", "cores": threads, "kind": "port",
                         "sample": f"{n_drr} DRR fwd+bwd(pose) of the {args.vol}^3->{args.det}^2 workload per step, "
                                   f"{len(times)} steps, C/OpenMP port of reference renderers.py (oracle/)"},
        "e2e": {"value": value, "unit": "DRRs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist

    from diffdrr_b200 import DRR, _lib, synthetic
    from diffdrr_b200.pose import convert
    from diffdrr_b200.renderers import _ptr, _stream, siddon_visits

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: diffdrr_b200 has no CPU path (use --impl reference for the CPU port)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, N, D = args.batch, args.det * args.det, args.vol

    # synthetic volume (performance is data independent) + module with the reference surface
    gen = torch.Generator(device=dev).manual_seed(0)
    vol = torch.rand(D, D, D, device=dev, generator=gen)
    subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
    subj.volume.affine = synthetic.make_affine(D)
    drr = DRR(subj, **synthetic.detector_kwargs(args.det), renderer="siddon").to(dev)
    drr.density = vol  # the 537 MB volume lives on the device only

    # this rank's poses (weak scaling: B per GPU); host copies live in pinned memory for the e2e leg
    rot_all, xyz_all = synthetic.make_poses(B * world, seed=0)
    rot_h = rot_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    xyz_h = xyz_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    with torch.no_grad():
        pose = convert(rot_h.to(dev), xyz_h.to(dev), parameterization="euler_angles", convention="ZXY")
        src, tgt = drr.detector(pose, None)
        raylen = (tgt - src).norm(dim=-1).unsqueeze(1).contiguous()
        src = drr.affine_inverse(src).reshape(B, 3).contiguous()
        tgt = drr.affine_inverse(tgt).contiguous()
    w = torch.rand(B, 1, args.det, args.det, device=dev, generator=gen)  # upstream gradient dLoss/dImage
    gout = w.reshape(B, N).contiguous()
    out = torch.empty(B, N, device=dev)
    g_src, g_tgt, g_len = torch.empty(B, 3, device=dev), torch.empty(B, N, 3, device=dev), torch.empty(B, N, device=dev)

    visits = siddon_visits((D, D, D), src, tgt)
    tot_visits = int(visits.sum().item())
    fwd_bytes = 4 * tot_visits + (4 + 16) * B * N            # voxels + out + (tgt, raylen)
    bwd_bytes = 4 * tot_visits + (4 + 16 + 12 + 4) * B * N   # voxels + gout + (tgt, raylen) + g_tgt + g_raylen

    stream = torch.cuda.current_stream()

    sens = torch.empty(B, N, 8, device=dev)
    sens_bytes = 4 * tot_visits + (16 + 4 + 32) * B * N      # voxels + (tgt, raylen) + out + 8 sensitivities per ray

    outs = [out, torch.empty_like(out)] if world > 1 else [out]
    gathered = [torch.empty(B * world, N, device=dev) for _ in range(2)] if world > 1 else None
    pending, state = [None, None], {"k": 0}

    def kernel_step(ev=None):
        """One training step at kernel level: image + per-ray sensitivities in ONE walk, then the elementwise backward."""
        out = outs[state["k"] & 1] if world > 1 else outs[0]
        if ev:
            ev[0].record(stream)
        _lib.check(lib.b200drr_siddon_fwd_sens_grid(_ptr(vol), D, D, D, _ptr(src), _ptr(tgt), _ptr(raylen), _ptr(out),
                                                    _ptr(sens), B, args.det, args.det, 0.5, 1e-8, 0, _stream()),
                   "siddon_fwd_sens_grid")
        if ev:
            ev[1].record(stream)
        _lib.check(lib.b200drr_siddon_bwd_sens(_ptr(sens), _ptr(gout), _ptr(g_src), _ptr(g_tgt), _ptr(g_len), B, N, 0,
                                               _stream()), "siddon_bwd_sens")
        if ev:
            ev[2].record(stream)
        if world > 1:
            # double-buffered: the gather of this step's image stack runs on NCCL's stream while the next step walks
            # (the next step writes the other buffer); a buffer is reused only after its gather has been waited for
            slot = state["k"] & 1
            if pending[slot] is not None:
                pending[slot].wait()
            pending[slot] = dist.all_gather_into_tensor(gathered[slot], outs[slot], async_op=True)
            state["k"] += 1

    def drain():
        for slot in (0, 1):
            if pending[slot] is not None:
                pending[slot].wait()
                pending[slot] = None

    def two_walk_step(ev):
        """The inference forward and the backward WALK (still the path when the volume needs gradients), for the record."""
        ev[0].record(stream)
        _lib.check(lib.b200drr_siddon_fwd_grid(_ptr(vol), D, D, D, _ptr(src), _ptr(tgt), _ptr(raylen), _ptr(out), B, args.det,
                                               args.det, 0.5, 1e-8, 0, _stream()), "siddon_fwd_grid")
        ev[1].record(stream)
        _lib.check(lib.b200drr_siddon_bwd_grid(_ptr(vol), D, D, D, _ptr(src), _ptr(tgt), _ptr(raylen), _ptr(gout), _ptr(g_src),
                                               _ptr(g_tgt), _ptr(g_len), None, B, args.det, args.det, 0.5, 1e-8, 0, 0,
                                               _stream()), "siddon_bwd_grid")
        ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernels-only leg (inputs resident in HBM) ------------------------------------------------------
    for _ in range(args.warmup):
        kernel_step()
    drain()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_start.record(stream)
        for k in range(args.steps):
            kernel_step(events[k])
        drain()  # the last gathers complete inside the timed region
        t_end.record(stream)
        barrier()
    ms_total = t_start.elapsed_time(t_end)
    sens_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in events]))
    sens_bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in events]))
    # not part of `value`: the plain forward kernel (inference) and the backward walk, timed the same way
    events2 = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        two_walk_step(events2[k])
    torch.cuda.synchronize()
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in events2[1:] or events2]))
    bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in events2[1:] or events2]))

    # ---- end-to-end leg through the public module, host buffers in and out -----------------------------
    img_h = torch.empty(B, 1, args.det, args.det).pin_memory()
    grad_h = torch.empty(2, B, 3).pin_memory()
    loss_h = torch.empty(1).pin_memory()

    def e2e_step():
        rot = rot_h.to(dev, non_blocking=True).requires_grad_(True)
        xyz = xyz_h.to(dev, non_blocking=True).requires_grad_(True)
        img = drr(rot, xyz, parameterization="euler_angles", convention="ZXY")
        loss = (img * w).sum()
        loss.backward()
        if world > 1:
            dist.all_gather_into_tensor(gathered[0], img.detach().reshape(B, N))
        img_h.copy_(img.detach(), non_blocking=True)
        grad_h[0].copy_(rot.grad, non_blocking=True)
        grad_h[1].copy_(xyz.grad, non_blocking=True)
        loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the result every step

    for _ in range(args.warmup):
        e2e_step()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e_end.record(stream)
    barrier()
    e2e_eager_ms_total = e_start.elapsed_time(e_end)

    # ---- the same end-to-end step captured ONCE in a CUDA graph and replayed (streams + graphs, no tracing compiler):
    # pinned-host pose -> H2D -> DRR(rot, xyz) -> loss -> backward -> D2H of image stack / loss / pose gradients.
    # The host buffers are re-read by every replay, so each step can carry new poses.  (Single GPU only: NCCL
    # collectives are kept out of the capture.)
    e2e_ms_total, e2e_mode = e2e_eager_ms_total, "eager"
    # N > 1: opt-in (B200DRR_BENCH_GRAPH_MULTI=1, not yet measured on a multi-GPU box): the per-rank step is replayed from
    # its graph and the gather of the image stack is issued after the replay, outside the capture.
    graph_multi = world > 1 and os.environ.get("B200DRR_BENCH_GRAPH_MULTI") == "1"
    if (world == 1 or graph_multi) and not args.no_graph:
        try:
            rot_d = torch.zeros(B, 3, device=dev, requires_grad=True)
            xyz_d = torch.zeros(B, 3, device=dev, requires_grad=True)
            rot_d.grad, xyz_d.grad = torch.zeros_like(rot_d), torch.zeros_like(xyz_d)

            copy_stream = torch.cuda.Stream()

            def graph_body():
                main = torch.cuda.current_stream()
                with torch.no_grad():
                    rot_d.copy_(rot_h, non_blocking=True)
                    xyz_d.copy_(xyz_h, non_blocking=True)
                    rot_d.grad.zero_()
                    xyz_d.grad.zero_()
                img = drr(rot_d, xyz_d, parameterization="euler_angles", convention="ZXY")
                # the 4 MB image stack goes back to the host on a second stream while loss + backward run
                # (a fork/join inside the captured graph)
                copy_stream.wait_stream(main)
                with torch.cuda.stream(copy_stream):
                    img_h.copy_(img.detach(), non_blocking=True)
                loss = (img * w).sum()
                loss.backward()
                main.wait_stream(copy_stream)
                grad_h[0].copy_(rot_d.grad, non_blocking=True)
                grad_h[1].copy_(xyz_d.grad, non_blocking=True)
                loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
                return img.detach()

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    graph_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            ref_img = img_h.clone()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                img_static = graph_body()

            def graph_step():
                graph.replay()
                if world > 1:
                    dist.all_gather_into_tensor(gathered[0], img_static.reshape(B, N))
                torch.cuda.current_stream().synchronize()  # the user reads the result every step

            for _ in range(args.warmup):
                graph_step()
            assert torch.allclose(img_h, ref_img, rtol=1e-4, atol=1e-3), "graph replay changed the images"
            g_start, g_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            g_start.record(stream)
            for _ in range(args.steps):
                graph_step()
            g_end.record(stream)
            barrier()
            e2e_ms_total, e2e_mode = g_start.elapsed_time(g_end), "cuda-graph replay of the public-API step"
        except Exception as exc:  # capture is an optimisation of the launch path only; never hide the eager number
            e2e_mode = f"eager (graph capture failed: {type(exc).__name__})"

    # ---- max over ranks -------------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, e2e_ms_total, fwd_ms, bwd_ms, e2e_eager_ms_total, sens_ms, sens_bwd_ms], device=dev,
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, fwd_ms, bwd_ms, e2e_eager_ms_total, sens_ms, sens_bwd_ms = (float(x) for x in stats.tolist())
    if rank != 0:
        return

    peak, peak_src = hbm_peak()
    total_drr = B * world * args.steps
    value = total_drr / (ms_total * 1e-3)
    e2e_value = total_drr / (e2e_ms_total * 1e-3)
    fwd_gbs = fwd_bytes / (fwd_ms * 1e-3) / 1e9
    bwd_gbs = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    sens_gbs = sens_bytes / (sens_ms * 1e-3) / 1e9
    dom = ("siddon_sens_slab_kernel", sens_gbs, sens_bytes, sens_ms)  # >95 % of the step
    line = {
        "metric": METRIC, "value": value, "unit": "DRRs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "volume": [D] * 3, "detector": [args.det] * 2, "batch_per_gpu": B,
                   "global_batch": B * world, "renderer": "siddon", "parallelism": f"pose-sharded dp{world}",
                   "l2": f"inputs > L2 ({4 * D ** 3 / 1e6:.0f} MB volume vs 126 MB L2); no explicit flush",
                   "mean_visits_per_ray": tot_visits / (B * N)},
        "e2e": {"value": e2e_value, "unit": "DRRs/s", "ms_per_step": e2e_ms_total / args.steps, "mode": e2e_mode,
                "eager_value": total_drr / (e2e_eager_ms_total * 1e-3), "eager_ms_per_step": e2e_eager_ms_total / args.steps,
                "h2d_bytes_per_step": int(rot_h.numel() + xyz_h.numel()) * 4,
                "d2h_bytes_per_step": int(img_h.numel() + grad_h.numel() + loss_h.numel()) * 4,
                "path": "DRR(rot, xyz) -> (img*w).sum().backward(); pinned host pose in; image stack + loss + pose grads out"},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom[1], "peak": peak, "unit": "GB/s", "frac": dom[1] / peak,
                     "traffic": NCU_TRAFFIC_BYTES.get(dom[0]) if (B, D, args.det) == (16, VOL, DET) else None, "algorithmic_bytes_per_launch": dom[2], "ms_per_launch": dom[3], "peak_source": peak_src},
        "step_breakdown_ms": {"siddon_sens_slab_kernel (image + sensitivities, one walk)": sens_ms,
                              "sens_bwd_kernel (elementwise backward)": sens_bwd_ms},
        "roofline_fwd": {"kernel": "siddon_fwd_slab_kernel", "achieved": fwd_gbs, "frac": fwd_gbs / peak, "ms_per_launch": fwd_ms,
                         "algorithmic_bytes_per_launch": fwd_bytes, "drr_per_s_fwd_only": B / (fwd_ms * 1e-3)},
        "roofline_bwd": {"kernel": "siddon_bwd_slab_kernel", "achieved": bwd_gbs, "frac": bwd_gbs / peak, "ms_per_launch": bwd_ms,
                         "algorithmic_bytes_per_launch": bwd_bytes,
                         "note": "two-walk backward; used when the volume needs gradients, not part of the timed step"},
        "clocks": clocks.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle

        threads = oracle.max_threads()
        times = cpu_oracle_time(D, args.det, 1, 3, threads)[1:]
        line["cpu_baseline"] = {"value": 1.0 / float(np.mean(times)), "unit": "DRRs/s", "cores": threads, "kind": "port",
                                "sample": f"1 DRR fwd+bwd(pose) of the same {D}^3->{args.det}^2 workload, mean of {len(times)} "
                                          "runs after 1 warm-up; C/OpenMP port of reference renderers.py (oracle/)"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
