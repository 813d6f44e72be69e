#!/usr/bin/env python
"""Add autograd gradients to the mask_to_channels goldens by running the UNMODIFIED reference on CPU.

Companion of make_golden.py (same shims, same inputs): reads the rays of `siddon_nc_b4_mask.npz` /
`trilinear_nc_b4_mask.npz`, the label volume `labels_nc.npz` and the volume `volumes.npz['nc']`, runs the reference
renderers with `mask=` (renderers.py:77-89, 242-252) in fp32 and fp64, and records the gradients of
loss = sum(w * img) for a stored weight w of shape (B, C, N) into `*_mask_grad.npz`.

    python tests/golden/make_golden_mask_grads.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from diffdrr.renderers import Siddon as RefSiddon, Trilinear as RefTrilinear  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

if __name__ == "__main__":
    vol = np.load(os.path.join(HERE, "volumes.npz"))["nc"]
    labels = np.load(os.path.join(HERE, "labels_nc.npz"))["labels"]
    for tag, cls, fkw in (("siddon_nc_b4_mask", RefSiddon, {}), ("trilinear_nc_b4_mask", RefTrilinear, dict(n_points=110))):
        g = np.load(os.path.join(HERE, tag + ".npz"))
        B, C, N = g["img_f64"].shape
        w = torch.rand(B, C, N, generator=torch.Generator().manual_seed(4321), dtype=torch.float64)
        rec = {"w": w.numpy()}
        for dt_tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            v = torch.from_numpy(vol).to(dt).requires_grad_(True)
            s = torch.from_numpy(g["source"]).to(dt).requires_grad_(True)
            t = torch.from_numpy(g["target"]).to(dt).requires_grad_(True)
            l = torch.from_numpy(g["raylen"]).to(dt).requires_grad_(True)
            img = cls()(v, s, t, l, mask=torch.from_numpy(labels).to(dt), **fkw)
            assert np.array_equal(img.detach().numpy(), g["img_" + dt_tag]), "forward differs from the stored golden"
            (img * w.to(dt)).sum().backward()
            for name, x in (("g_volume", v), ("g_source", s), ("g_target", t), ("g_raylen", l)):
                rec[f"{name}_{dt_tag}"] = x.grad.numpy()
        path = os.path.join(HERE, tag + "_grad.npz")
        np.savez_compressed(path, **rec)
        print(tag, {k: (v.shape, float(np.abs(v).max())) for k, v in rec.items() if k.endswith("f64")},
              f"{os.path.getsize(path) / 1024:.0f} KiB")
