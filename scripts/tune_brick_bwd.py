#!/usr/bin/env python
"""Volume gradient of the Siddon render (reconstruction path): brick kernel in scatter mode (b200drr_siddon_bwd_vol_brick:
shared-memory accumulation, one TMA store per brick) against the slab-major walk with one global atomic per voxel visit
(b200drr_siddon_bwd_grid with g_vol), 512^3 (VOL) -> 256^2 (DET), B poses; parity against each other and, for a small case,
nothing else (tests/ do that)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffdrr_b200 import DRR, _lib, synthetic  # noqa: E402
from diffdrr_b200.renderers import _ptr, _stream, siddon_visits  # noqa: E402

D = int(os.environ.get("VOL", 512))
dev = torch.device("cuda:0")
lib = _lib.load()
peak = bench.hbm_peak()[0]
vol = torch.as_tensor(synthetic.make_volume(D, "rand", seed=0)).to(dev)
subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
subj.volume.affine = synthetic.make_affine(D)
for H, B in [(int(x.split("x")[0]), int(x.split("x")[1])) for x in os.environ.get("CASES", "256x16,256x1,256x4,512x8").split(",")]:
    drr = DRR(subj, **synthetic.detector_kwargs(H)).to(dev)
    rot, xyz = synthetic.make_poses(max(B, 2), seed=0)
    s, t, l = bench._device_rays(drr, rot[:B], xyz[:B], dev)
    N = H * H
    gout = torch.rand(B, N, device=dev)
    visits = int(siddon_visits((D, D, D), s, t).sum().item())
    g_a, g_b = torch.zeros_like(vol), torch.empty_like(vol)
    ws = torch.empty(int(lib.b200drr_siddon_brick_workspace_bytes(B, H, H)), dtype=torch.uint8, device=dev)

    def slab():
        g_a.zero_()
        _lib.check(lib.b200drr_siddon_bwd_grid(_ptr(vol), D, D, D, _ptr(s), _ptr(t), _ptr(l), _ptr(gout), None, None, None, _ptr(g_a),
                                               B, H, H, 0.5, 1e-8, 0, 0, _stream()), "bwd_grid")

    def brick():
        _lib.check(lib.b200drr_siddon_bwd_vol_brick(_ptr(gout), D, D, D, _ptr(s), _ptr(t), _ptr(l), None, None, None, None, _ptr(g_b),
                                                    ctypes.c_void_p(ws.data_ptr()), ws.numel(), B, H, H, 0.5, 1e-8, _stream()),
                   "bwd_vol_brick")

    ta = float(np.median(bench._time_events(slab, 5, warmup=2)))
    tb = float(np.median(bench._time_events(brick, 5, warmup=2)))
    err = float((g_a - g_b).abs().max() / g_a.abs().max())
    alg = 8 * visits + 4 * D ** 3  # read-modify-write of every visited accumulator + one store of the volume
    print(f"{D}^3 -> {H}^2 x {B:2d} poses: slab-major global atomics {ta:7.3f} ms | brick scatter {tb:7.3f} ms ({ta / tb:4.2f}x) "
          f"= {B / tb * 1e3:8.1f} DRR/s, {4 * visits / tb * 1e-6:7.1f} GB/s of visited voxels ({4 * visits / tb * 1e-6 / peak * 100:4.1f} % of HBM peak)"
          f"  maxdiff/max {err:.1e}  nan {bool(torch.isnan(g_b).any())}")
