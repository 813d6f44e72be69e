#!/usr/bin/env python
"""Records NormalizedCrossCorrelation2d values + gradients from the UNMODIFIED reference (diffdrr/metrics.py:21-44), incl.
the patch mode (metrics.py:16-19,30).  metrics.py also imports kornia / torchvision for OTHER losses; both are stubbed here
(never in the product).  Run in the build container: python tests/golden/make_golden_ncc.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "_refshim"))
sys.path.insert(0, "/root/reference")
for name in ("kornia", "kornia.enhance", "kornia.enhance.histogram"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["kornia.enhance.histogram"].marginal_pdf = sys.modules["kornia.enhance.histogram"].joint_pdf = None
if "torchvision" not in sys.modules:
    try:
        import torchvision  # noqa: F401
    except Exception:
        for name in ("torchvision", "torchvision.transforms", "torchvision.transforms.functional"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["torchvision.transforms.functional"].gaussian_blur = None
from diffdrr.metrics import NormalizedCrossCorrelation2d  # noqa: E402  (the reference's own class)

g = torch.Generator().manual_seed(0)
x1 = torch.rand(3, 1, 24, 20, generator=g)
x2 = (0.6 * x1 + 0.4 * torch.rand(3, 1, 24, 20, generator=g)).contiguous()
rec = {"x1": x1.numpy(), "x2": x2.numpy()}
for tag, patch in (("full", None), ("patch5", 5)):
    for dt, suf in ((torch.float32, "f32"), (torch.float64, "f64")):
        a = x1.to(dt)
        b = x2.to(dt).clone().detach().requires_grad_(True)
        score = NormalizedCrossCorrelation2d(patch_size=patch)(a, b)
        w = torch.tensor([1.0, -2.0, 0.5], dtype=dt)
        (score * w).sum().backward()
        rec[f"{tag}_score_{suf}"] = score.detach().numpy()
        rec[f"{tag}_grad_x2_{suf}"] = b.grad.numpy()
np.savez_compressed(os.path.join(HERE, "ncc_reference.npz"), **rec)
print({k: v.shape for k, v in rec.items()})
