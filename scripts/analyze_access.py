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
    out = np.zeros(5)
    s, t = src, tgt
    if poses is not None:
        s, t = np.ascontiguousarray(src[poses]), np.ascontiguousarray(tgt[poses])
    lib.emu_warp_sectors(D, D, D, s.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p), len(s), H, H, WX, WY, slab,
                         ctypes.c_long(strides[0]), ctypes.c_long(strides[1]), ctypes.c_long(strides[2]), ctypes.c_float(0.5), ctypes.c_float(1e-8), 37, out.ctypes.data_as(ctypes.c_void_p))
    return out
layouts = {"linear [D0][D1][D2]": (D * D, D, 1), "brick 4x4x2": (0, 0, -1), "brick 2x4x4": (0, 0, -2), "brick 4x2x4": (0, 0, -3)}
for name, st in layouts.items():
    for (wx, wy) in [(8, 4), (4, 8)]:
        o = run(wx, wy, 32, st)
        print(f"{name:20s} warp {wx}x{wy}: lanes/step {o[1]/o[0]:5.1f}  sectors/step {o[2]/o[0]:5.1f}  lines/step {o[3]/o[0]:5.1f}  "
              f"sectors/visit {o[2]/o[1]:.3f}  distinct-sectors-per-visit over the item {o[4]/o[1]:.3f}")
