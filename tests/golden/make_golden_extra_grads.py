#!/usr/bin/env python
"""Autograd gradients of the UNMODIFIED reference for the option cases that make_golden.py records forward-only:
`reducefn="max"` (both renderers) and Siddon `align_corners=True`.  Same inputs and weight image `w` as the stored cases;
output `<case>_grad.npz` with g_volume / g_source / g_target / g_raylen in fp32 and fp64.

    python tests/golden/make_golden_extra_grads.py
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

CASES = (
    ("siddon_nc_b4_max", RefSiddon, dict(reducefn="max"), {}),
    ("siddon_nc_b4_ac", RefSiddon, {}, dict(align_corners=True)),
    ("trilinear_nc_b4_max", RefTrilinear, dict(reducefn="max"), None),
)

if __name__ == "__main__":
    vol = np.load(os.path.join(HERE, "volumes.npz"))["nc"]
    for tag, cls, ctor, fkw in CASES:
        g = np.load(os.path.join(HERE, tag + ".npz"))
        if fkw is None:
            fkw = dict(n_points=96)  # make_golden.py: trilinear_nc_b4_max
        rec = {}
        for dt_tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            v = torch.from_numpy(vol).to(dt).requires_grad_(True)
            s = torch.from_numpy(g["source"]).to(dt).requires_grad_(True)
            t = torch.from_numpy(g["target"]).to(dt).requires_grad_(True)
            l = torch.from_numpy(g["raylen"]).to(dt).requires_grad_(True)
            img = cls(**ctor)(v, s, t, l, **fkw)
            assert np.array_equal(img.detach().numpy(), g["img_" + dt_tag]), f"{tag}: forward differs from the stored golden"
            (img * torch.from_numpy(g["w"]).to(dt)).sum().backward()
            for name, x in (("g_volume", v), ("g_source", s), ("g_target", t), ("g_raylen", l)):
                rec[f"{name}_{dt_tag}"] = x.grad.numpy()
        path = os.path.join(HERE, tag + "_grad.npz")
        np.savez_compressed(path, **rec)
        print(tag, {k: float(np.abs(v).max()) for k, v in rec.items() if k.endswith("f64")}, f"{os.path.getsize(path) / 1024:.0f} KiB")

    # Siddon with mode="bilinear" (renderers.py:18,66: trilinear sampling at the segment midpoints), with and without
    # stop_gradients_through_grid_sample; rays and weights of siddon_nc_b4; full records (no forward golden existed)
    g = np.load(os.path.join(HERE, "siddon_nc_b4.npz"))
    for tag, ctor in (("siddon_nc_b4_bilinear", dict(mode="bilinear")),
                      ("siddon_nc_b4_bilinear_stopgrad", dict(mode="bilinear", stop_gradients_through_grid_sample=True))):
        rec = {k: g[k] for k in ("source", "target", "raylen", "w")}
        rec["volume_key"] = np.str_("nc")
        for dt_tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            v = torch.from_numpy(vol).to(dt).requires_grad_(True)
            s = torch.from_numpy(g["source"]).to(dt).requires_grad_(True)
            t = torch.from_numpy(g["target"]).to(dt).requires_grad_(True)
            l = torch.from_numpy(g["raylen"]).to(dt).requires_grad_(True)
            img = RefSiddon(**ctor)(v, s, t, l)
            rec["img_" + dt_tag] = img.detach().numpy()
            (img * torch.from_numpy(g["w"]).to(dt)).sum().backward()
            for name, x in (("g_volume", v), ("g_source", s), ("g_target", t), ("g_raylen", l)):
                if x.grad is not None:
                    rec[f"{name}_{dt_tag}"] = x.grad.numpy()
        path = os.path.join(HERE, tag + ".npz")
        np.savez_compressed(path, **rec)
        print(tag, {k: float(np.abs(v).max()) for k, v in rec.items() if k.endswith("f64")}, f"{os.path.getsize(path) / 1024:.0f} KiB")
