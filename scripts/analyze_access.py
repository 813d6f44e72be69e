#!/usr/bin/env python
"""CPU-side tuning aid: distinct sectors/lines per warp-wide gather for warp shapes / layouts (uses tests/hostemu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from hostemu import emu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_host_rays

D, H, B = 512, 256, 16
src, tgt, _ = make_host_rays(D, H, B, seed=0)
src = np.ascontiguousarray(src.reshape(B, 3), np.float32); tgt = np.ascontiguousarray(tgt, np.float32)
lib = emu.lib()
def run(WX, WY, slab, strides, poses=None):
    out = np.zeros(4)
    s, t = src, tgt
    if poses is not None:
        s, t = np.ascontiguousarray(src[poses]), np.ascontiguousarray(tgt[poses])
    lib.emu_warp_sectors(D, D, D, s.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p), len(s), H, H, WX, WY, slab,
                         ctypes.c_long(strides[0]), ctypes.c_long(strides[1]), ctypes.c_long(strides[2]), ctypes.c_float(0.5), ctypes.c_float(1e-8), 37, out.ctypes.data_as(ctypes.c_void_p))
    return out
orig = (D * D, D, 1)
layouts = {"orig(fast=2)": (D * D, D, 1), "T02(fast=0)": (1, D, D * D), "T12(fast=1)": (D * D, 1, D)}
for name, st in layouts.items():
    for (wx, wy) in [(8, 4), (4, 8), (16, 2), (2, 16), (32, 1), (1, 32)]:
        o = run(wx, wy, 32, st)
        print(f"{name:14s} warp {wx:2d}x{wy:<2d}: lanes/step {o[1]/o[0]:5.1f}  sectors/step {o[2]/o[0]:5.1f}  lines/step {o[3]/o[0]:5.1f}  sectors/visit {o[2]/o[1]:.3f}")
print("per pose, orig layout, 8x4 / best-of-shapes / best over layouts:")
for b in range(B):
    res = {}
    for name, st in layouts.items():
        for (wx, wy) in [(8, 4), (4, 8), (16, 2), (2, 16)]:
            o = run(wx, wy, 32, st, [b]); res[(name, wx, wy)] = o[2] / o[1]
    d = tgt[b, H * H // 2 + H // 2] - src[b]
    d = d / np.linalg.norm(d)
    best = min(res, key=res.get)
    print(f" pose {b:2d} dir {np.round(d,2)}  8x4 orig {res[('orig(fast=2)',8,4)]:.3f}  best orig {min(v for k,v in res.items() if k[0].startswith('orig')):.3f}  best all {res[best]:.3f} {best}")
