#!/usr/bin/env python
"""mask_to_channels cost (reference renderers.py:77-89 / 242-252; the reference publishes +54 % for it): Siddon and trilinear
forward with a label volume, row-ordered threads (b200drr_*_fwd_mask) vs tile-ordered (b200drr_*_fwd_mask_grid) vs the plain
un-masked renderers, 512^3 (or VOL) -> 256^2, 16 poses, 8 labels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffdrr_b200 import DRR, Siddon, Trilinear, synthetic  # noqa: E402

D, H, B = int(os.environ.get("VOL", 512)), int(os.environ.get("DET", 256)), int(os.environ.get("B", 16))
dev = torch.device("cuda:0")
vol = torch.as_tensor(synthetic.make_volume(D, "rand", seed=0)).to(dev)
x = torch.linspace(-1, 1, D, device=dev)
lab = ((x[:, None, None] > 0).float() + 2 * (x[None, :, None] > 0).float() + 4 * (x[None, None, :] > 0).float()).contiguous()
subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
subj.volume.affine = synthetic.make_affine(D)
drr = DRR(subj, **synthetic.detector_kwargs(H)).to(dev)
rot, xyz = synthetic.make_poses(B, seed=0)
s, t, l = bench._device_rays(drr, rot, xyz, dev)
s, l = s.reshape(B, 1, 3), l.reshape(B, 1, H * H)


def timed(fn, n=5):
    ts = bench._time_events(fn, n, warmup=2)
    return float(np.median(ts))


with torch.no_grad():
    for name, mod, kw in (("siddon", Siddon(), {}), ("trilinear", Trilinear(), dict(n_points=500))):
        res = {}
        for tag, grid in (("rows", None), ("tiles", (H, H))):
            mod.detector_shape = grid
            res["plain " + tag] = timed(lambda: mod(vol, s, t, l, **kw))
            res["mask " + tag] = timed(lambda: mod(vol, s, t, l, mask=lab, **kw))
        # backward of the masked render (ray gradients; label routing of the upstream gradient), rows vs tiles
        g = torch.rand(B, 8, H * H, device=dev)
        for tag, grid in (("rows", None), ("tiles", (H, H))):
            mod.detector_shape = grid
            with torch.enable_grad():
                tt = t.detach().clone().requires_grad_(True)
                out = mod(vol, s, tt, l, mask=lab, **kw)

                def bwd():
                    tt.grad = None
                    out.backward(g, retain_graph=True)

                res["mask bwd " + tag] = timed(bwd, 3)
        mod.detector_shape = None
        a = mod(vol, s, t, l, mask=lab, **kw)
        mod.detector_shape = (H, H)
        b = mod(vol, s, t, l, mask=lab, **kw)
        full = mod(vol, s, t, l, **kw)
        print(f"{name:9s} {D}^3 -> {H}^2 x {B} poses, 8 labels: " + "  ".join(f"{k} {v:.3f} ms" for k, v in res.items())
              + f"  | mask/plain (tiles) {res['mask tiles'] / res['plain tiles']:.2f}x"
              + f"  tiles vs rows maxdiff {float((a - b).abs().max()):.1e}  sum over channels vs plain relerr "
              + f"{float((b.sum(1, keepdim=True) - full).abs().max() / full.abs().max()):.1e}")
