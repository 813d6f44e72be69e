#!/usr/bin/env python
"""CPU stress of the cut conventions: dyadic geometry makes every plane alpha exact in fp32, so almost every
crossing of every ray ties with another one; the sensitivities walk cut into major-axis pieces (and into slabs along axis 0) must
give the un-cut walk's per-ray gradients.  Runs the product's device math through tests/hostemu (test infrastructure).
usage: scripts/stress_cut_ties.py [seed] [trials]   -> prints the rays that differ and `bad N of M`"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hostemu import emu
from conftest import relerr
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0; tot = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    dims = tuple(int(x) for x in rng.integers(8, 40, 3))
    g = np.meshgrid(*[np.linspace(-1, 1, d) for d in dims], indexing="ij")
    vol = (np.exp(-(g[0]**2 + (g[1]-0.2)**2 + g[2]**2) / 0.5) + 0.3 * rng.random(dims)).astype(np.float32)
    n = 300
    # dyadic geometry: alphas are exact in fp32, ties are EXACT
    d = rng.choice([32.0, 64.0, 128.0], size=(n, 3)) * rng.choice([-1.0, 1.0], size=(n, 3))
    through = np.stack([rng.integers(0, dims[a] + 1, n).astype(np.float64) for a in range(3)], 1) - 0.5   # a lattice point (plane space - shift)
    through += rng.choice([0.0, 0.0, 0.25, 0.5], size=(n, 1)) * rng.choice([0.0, 1.0], size=(n, 3))
    k = rng.integers(1, 4, size=(n, 1)).astype(np.float64)
    src_all = through - d * k * 0.25 - 0.0
    # one source per pose in the API: loop over rays as separate "poses" in blocks sharing a source -> use B = n, N = 1
    s = src_all.astype(np.float32).reshape(n, 1, 3); t = (src_all + d * 2.0).astype(np.float32).reshape(n, 1, 3)
    l = np.linalg.norm(t - s, axis=-1).reshape(n, 1, 1).astype(np.float32)
    w = np.ones((n, 1, 1), np.float32)
    K = int(rng.integers(2, 45))
    a = emu.siddon_sens(vol, s, t, l, w, slab=-K); u = emu.siddon_sens(vol, s, t, l, w, slab=0)
    sl = emu.siddon_sens(vol, s, t, l, w, slab=int(rng.integers(2, 9)))
    scale = max(np.abs(u["g_target"]).max(), 1e-20)
    dt = np.abs(a["g_target"] - u["g_target"]).max(-1)[:, 0] / scale
    dsl = np.abs(sl["g_target"] - u["g_target"]).max(-1)[:, 0] / scale
    di = np.abs(a["img"] - u["img"])[:, 0, 0] / max(np.abs(u["img"]).max(), 1e-20)
    nb = int((dt > 1e-4).sum()); tot += n; bad += nb
    if nb or (di > 1e-5).any() or (dsl > 1e-4).any():
        i = int(np.argmax(dt))
        print(f"trial {trial} dims {dims} K {K}: MAJ-cut bad rays {nb}/{n} (max {dt.max():.2e}), slab-cut bad {(dsl > 1e-4).sum()} (max {dsl.max():.2e}), img max {di.max():.2e}; worst ray src {s[i,0]} d {d[i]}")
print("bad", bad, "of", tot)
