#!/usr/bin/env python
"""Per-ray cost of sub-sampled ray sets: locality-ordered slab-major kernels vs one thread per ray vs the full grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

os.environ["SMALL"] = "0"
from diffdrr_b200 import Siddon, renderers  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("tb", os.path.join(os.path.dirname(__file__), "tune_brick.py"))
os.environ["BVARIANTS"] = "0"
tb = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0]]
try:
    spec.loader.exec_module(tb)
except ValueError:
    pass
dev, B, H, dims = tb.dev, tb.B, tb.H, tb.dims
vol, src, tgt, raylen = tb.vol, tb.src, tb.tgt, tb.raylen
N = H * H
full = tb.timeit(lambda: tb.grid_call(vol, dims, src, tgt, raylen, torch.empty(B, N, device=dev), B, H, H))
print(f"full grid {H}^2, {B} poses: {full:.3f} ms = {full * 1e6 / (B * N):.2f} ns/ray")
for frac in (0.5, 0.25, 0.1):
    sel = torch.randperm(N, device=dev, generator=torch.Generator(device=dev).manual_seed(0))[: int(N * frac)].sort().values
    tg, ln = tgt[:, sel].contiguous(), raylen[:, sel].contiguous().reshape(B, 1, -1)
    s3 = src.reshape(B, 1, 3)
    for name, thr in (("sorted slab-major", 1), ("one thread per ray", 10 ** 9)):
        renderers._SORT_MIN_RAYS = thr
        with torch.no_grad():
            ms = tb.timeit(lambda: Siddon()(vol, s3, tg, ln))
        print(f"  {frac * 100:4.0f}% sub-sample, {name:20s}: {ms:.3f} ms = {ms * 1e6 / (B * len(sel)):.2f} ns/ray  ({ms * 1e6 / (B * len(sel)) / (full * 1e6 / (B * N)):.2f}x the full-grid per-ray cost)")
    st, tt = s3.clone().requires_grad_(True), tg.clone().requires_grad_(True)
    for name, thr in (("sorted slab-major", 1), ("one thread per ray", 10 ** 9)):
        renderers._SORT_MIN_RAYS = thr
        def step():
            o = Siddon()(vol, st, tt, ln)
            o.sum().backward()
        ms = tb.timeit(step)
        print(f"  {frac * 100:4.0f}% sub-sample fwd+bwd(pose), {name:20s}: {ms:.3f} ms")

# ---- sub-sampled detector through the module (inference): brick-major kernel with a pixel -> ray map vs the sorted path
from diffdrr_b200 import DRR, synthetic  # noqa: E402
for frac in (0.5, 0.25, 0.1):
    subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
    subj.volume.affine = synthetic.make_affine(dims)
    torch.manual_seed(0)
    drr = DRR(subj, **synthetic.detector_kwargs(H), p_subsample=frac).to(dev)
    drr.density = vol
    rot, xyz = synthetic.make_poses(B, seed=0)
    rot, xyz = rot.to(dev), xyz.to(dev)
    for name, mb in (("brick-major + pixel map", 2), ("sorted slab-major", 10 ** 9)):
        renderers._BRICK_MIN_BATCH = mb
        renderers._SORT_MIN_RAYS = 1
        with torch.no_grad():
            ms = tb.timeit(lambda: drr(rot, xyz, parameterization="euler_angles", convention="ZXY"))
        n_sub = int(N * frac)
        print(f"  DRR(p_subsample={frac}) inference, {name:24s}: {ms:.3f} ms (module call incl. pose algebra) = {ms * 1e6 / (B * n_sub):.2f} ns/ray")

# ---- the subset kernel alone (C ABI, rays resident): per-ray cost vs the full grid
from diffdrr_b200 import _lib  # noqa: E402
from diffdrr_b200.renderers import _ptr, _stream  # noqa: E402
lib = _lib.load()
ws_full = torch.empty(lib.b200drr_siddon_brick_workspace_bytes(B, H, H), dtype=torch.uint8, device=dev)
out_full = torch.empty(B, N, device=dev)
t_full = tb.timeit(lambda: tb.brick_call(vol, dims, src, tgt, raylen, out_full, ws_full, B, H, H, 0))
print(f"brick-major full grid: {t_full:.3f} ms = {t_full * 1e6 / (B * N):.2f} ns/ray")
corners = tgt[:, [0, H - 1, (H - 1) * H], :].contiguous()
for frac in (0.5, 0.25, 0.1):
    sel = torch.randperm(N, device=dev, generator=torch.Generator(device=dev).manual_seed(0))[: int(N * frac)]
    ns = len(sel)
    pix = torch.full((N,), -1, dtype=torch.int32, device=dev)
    pix[sel] = torch.arange(ns, dtype=torch.int32, device=dev)
    tg, ln = tgt[:, sel].contiguous(), raylen[:, sel].contiguous()
    o = torch.empty(B, ns, device=dev)
    ws = torch.empty(lib.b200drr_siddon_brick_workspace_bytes(B, 1, ns), dtype=torch.uint8, device=dev)
    def call():
        _lib.check(lib.b200drr_siddon_fwd_brick_subset(_ptr(vol), *dims, _ptr(src), _ptr(tg), _ptr(ln), pix.data_ptr(), _ptr(corners), _ptr(o),
                                                       ws.data_ptr(), ws.numel(), B, H, H, ns, 0.5, 1e-8, 0, _stream()), "subset")
    ms = tb.timeit(call)
    err = float((o - out_full[:, sel]).abs().max() / out_full.abs().max())
    print(f"  brick-major {frac * 100:4.0f}% sub-sample (kernel only): {ms:.3f} ms = {ms * 1e6 / (B * ns):.2f} ns/ray ({ms * 1e6 / (B * ns) / (t_full * 1e6 / (B * N)):.2f}x the brick full-grid per-ray cost, {ms * 1e6 / (B * ns) / (full * 1e6 / (B * N)):.2f}x the slab-major full grid's)  maxdiff {err:.1e}")
