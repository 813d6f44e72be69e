#!/usr/bin/env python
"""Images and autograd gradients of the UNMODIFIED reference with a CALLABLE `reducefn` (renderers.py:175-183): the callable
receives the (B, N, M-1) per-segment (Siddon) / (B, N, n_points) per-sample (trilinear) tensor.  Rays and weight image of the
stored siddon_nc_b4 / trilinear_nc_b4 cases; reducers from tests/callable_reducers.py.  Output `<case>_callable.npz`.

    python tests/golden/make_golden_callable.py
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

from diffdrr.renderers import Siddon as RefSiddon, Trilinear as RefTrilinear  # noqa: E402
from callable_reducers import REDUCERS  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

if __name__ == "__main__":
    vol = np.load(os.path.join(HERE, "volumes.npz"))["nc"]
    for tag, cls, fkw in (("siddon_nc_b4", RefSiddon, {}), ("trilinear_nc_b4", RefTrilinear, dict(n_points=160))):
        g = np.load(os.path.join(HERE, tag + ".npz"))
        rec = {k: g[k] for k in ("source", "target", "raylen", "w")}
        rec["volume_key"] = np.str_("nc")
        for rname, fn in REDUCERS.items():
            for dt_tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                v = torch.from_numpy(vol).to(dt).requires_grad_(True)
                s = torch.from_numpy(g["source"]).to(dt).requires_grad_(True)
                t = torch.from_numpy(g["target"]).to(dt).requires_grad_(True)
                l = torch.from_numpy(g["raylen"]).to(dt).requires_grad_(True)
                img = cls(reducefn=fn)(v, s, t, l, **fkw)
                rec[f"{rname}_img_{dt_tag}"] = img.detach().numpy()
                (img * torch.from_numpy(g["w"]).to(dt)).sum().backward()
                for name, x in (("g_volume", v), ("g_source", s), ("g_target", t), ("g_raylen", l)):
                    rec[f"{rname}_{name}_{dt_tag}"] = x.grad.numpy()
        path = os.path.join(HERE, tag + "_callable.npz")
        np.savez_compressed(path, **rec)
        print(tag, {k: float(np.abs(v).max()) for k, v in rec.items() if k.endswith("f64")}, f"{os.path.getsize(path) / 1024:.0f} KiB")
