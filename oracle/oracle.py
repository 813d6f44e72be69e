"""ctypes front-end of oracle/liboracle.so (the plain-C restatement of reference renderers.py).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs -- never by the product package `diffdrr_b200/`.  All arrays are numpy,
C-contiguous; `dtype` selects the fp32 (reference-literal) or fp64 (ground-truth) build of each function.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (Makefile recipe); building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("drr_oracle.c", "drr_oracle_impl.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def _prep(dtype, *arrays):
    return [None if a is None else np.ascontiguousarray(a, dtype=dtype) for a in arrays]


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _real(dtype):
    return ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double


def _suf(dtype):
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def siddon_fwd(vol, src, tgt, raylen, voxel_shift=0.5, eps=1e-8, reduce="sum", align_corners=False, dtype=np.float32):
    """vol (D0,D1,D2); src (B,1,3)|(B,3); tgt (B,N,3); raylen (B,1,N)|(B,N) -> (B,1,N)."""
    vol, src, tgt, raylen = _prep(dtype, vol, src, tgt, raylen)
    B, N = tgt.shape[0], tgt.shape[1]
    out = np.empty((B, 1, N), dtype=dtype)
    R = _real(dtype)
    getattr(lib(), "oracle_siddon_fwd_" + _suf(dtype))(
        _p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
        ctypes.c_long(N), R(voxel_shift), R(eps), ctypes.c_int({"sum": 0, "max": 1}[reduce]),
        ctypes.c_int(bool(align_corners)))
    return out


def siddon_fwd_mask(vol, mask, src, tgt, raylen, n_channels, voxel_shift=0.5, eps=1e-8, align_corners=False, dtype=np.float32):
    """mask_to_channels rendering: (B, C, N)."""
    vol, mask, src, tgt, raylen = _prep(dtype, vol, mask, src, tgt, raylen)
    B, N = tgt.shape[0], tgt.shape[1]
    out = np.empty((B, n_channels, N), dtype=dtype)
    R = _real(dtype)
    getattr(lib(), "oracle_siddon_fwd_mask_" + _suf(dtype))(
        _p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
        ctypes.c_long(N), ctypes.c_int(n_channels), R(voxel_shift), R(eps), ctypes.c_int(bool(align_corners)))
    return out


def trilinear_fwd_mask(vol, mask, src, tgt, raylen, n_channels, n_points=500, alphamin=None, alphamax=None, voxel_shift=0.5,
                       eps=1e-8, align_corners=False, dtype=np.float32):
    vol, mask, src, tgt, raylen = _prep(dtype, vol, mask, src, tgt, raylen)
    B, N = tgt.shape[0], tgt.shape[1]
    if alphamin is None or alphamax is None:
        alphamin, alphamax = alpha_minmax(vol.shape, src, tgt, voxel_shift, eps, dtype)
    out = np.empty((B, n_channels, N), dtype=dtype)
    R = _real(dtype)
    getattr(lib(), "oracle_trilinear_fwd_mask_" + _suf(dtype))(
        _p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
        ctypes.c_long(N), ctypes.c_int(n_channels), R(voxel_shift), R(eps), ctypes.c_int(n_points), R(alphamin),
        R(alphamax), ctypes.c_int(bool(align_corners)))
    return out


def siddon_bwd(vol, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False, align_corners=False,
               want_vol=True, dtype=np.float64, reduce="sum"):
    """Returns dict(g_source (B,1,3), g_target (B,N,3), g_raylen (B,1,N), g_volume (D0,D1,D2)|None)."""
    vol, src, tgt, raylen, gout = _prep(dtype, vol, src, tgt, raylen, gout)
    B, N = tgt.shape[0], tgt.shape[1]
    g_src = np.zeros((B, 1, 3), dtype=dtype)
    g_tgt = np.zeros((B, N, 3), dtype=dtype)
    g_len = np.zeros((B, 1, N), dtype=dtype)
    g_vol = np.zeros(vol.shape, dtype=dtype) if (want_vol and not stop_grad) else None
    R = _real(dtype)
    getattr(lib(), ("oracle_siddon_bwd_max_" if reduce == "max" else "oracle_siddon_bwd_") + _suf(dtype))(
        _p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src), _p(g_tgt),
        _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N), R(voxel_shift), R(eps), ctypes.c_int(bool(stop_grad)),
        ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def alpha_minmax(vol_shape, src, tgt, voxel_shift=0.5, eps=1e-8, dtype=np.float32):
    src, tgt = _prep(dtype, src, tgt)
    B, N = tgt.shape[0], tgt.shape[1]
    R = _real(dtype)
    lo, hi = R(0), R(0)
    getattr(lib(), "oracle_alpha_minmax_" + _suf(dtype))(
        _p(src), _p(tgt), *map(ctypes.c_int, vol_shape), ctypes.c_int(B), ctypes.c_long(N), R(voxel_shift), R(eps),
        ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


def trilinear_fwd(vol, src, tgt, raylen, n_points=500, alphamin=None, alphamax=None, voxel_shift=0.5, eps=1e-8,
                  reduce="sum", align_corners=False, dtype=np.float32):
    vol, src, tgt, raylen = _prep(dtype, vol, src, tgt, raylen)
    B, N = tgt.shape[0], tgt.shape[1]
    if alphamin is None or alphamax is None:
        alphamin, alphamax = alpha_minmax(vol.shape, src, tgt, voxel_shift, eps, dtype)
    out = np.empty((B, 1, N), dtype=dtype)
    R = _real(dtype)
    getattr(lib(), "oracle_trilinear_fwd_" + _suf(dtype))(
        _p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
        ctypes.c_long(N), R(voxel_shift), R(eps), ctypes.c_int(n_points), R(alphamin), R(alphamax),
        ctypes.c_int({"sum": 0, "max": 1}[reduce]), ctypes.c_int(bool(align_corners)))
    return out


def trilinear_bwd(vol, src, tgt, raylen, gout, n_points=500, alphamin=None, alphamax=None, voxel_shift=0.5, eps=1e-8,
                  align_corners=False, want_vol=True, dtype=np.float64, reduce="sum"):
    """Gradients for FIXED alphamin/alphamax plus the partials g_alphamin/g_alphamax (scalars)."""
    vol, src, tgt, raylen, gout = _prep(dtype, vol, src, tgt, raylen, gout)
    B, N = tgt.shape[0], tgt.shape[1]
    if alphamin is None or alphamax is None:
        alphamin, alphamax = alpha_minmax(vol.shape, src, tgt, voxel_shift, eps, dtype)
    g_src = np.zeros((B, 1, 3), dtype=dtype)
    g_tgt = np.zeros((B, N, 3), dtype=dtype)
    g_len = np.zeros((B, 1, N), dtype=dtype)
    g_vol = np.zeros(vol.shape, dtype=dtype) if want_vol else None
    R = _real(dtype)
    ga0, ga1 = R(0), R(0)
    getattr(lib(), ("oracle_trilinear_bwd_max_" if reduce == "max" else "oracle_trilinear_bwd_") + _suf(dtype))(
        _p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src), _p(g_tgt),
        _p(g_len), _p(g_vol), ctypes.byref(ga0), ctypes.byref(ga1), ctypes.c_int(B), ctypes.c_long(N), R(voxel_shift),
        R(eps), ctypes.c_int(n_points), R(alphamin), R(alphamax), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol, g_alphamin=ga0.value,
                g_alphamax=ga1.value)


def siddon_bwd_mask(vol, mask, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False, align_corners=False,
                    want_vol=True, dtype=np.float64):
    """Backward of siddon_fwd_mask for gout (B,C,N); same outputs as siddon_bwd."""
    vol, mask, src, tgt, raylen, gout = _prep(dtype, vol, mask, src, tgt, raylen, gout)
    B, N, C = tgt.shape[0], tgt.shape[1], gout.shape[1]
    g_src, g_tgt = np.zeros((B, 1, 3), dtype=dtype), np.zeros((B, N, 3), dtype=dtype)
    g_len = np.zeros((B, 1, N), dtype=dtype)
    g_vol = np.zeros(vol.shape, dtype=dtype) if (want_vol and not stop_grad) else None
    R = _real(dtype)
    getattr(lib(), "oracle_siddon_bwd_mask_" + _suf(dtype))(
        _p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src), _p(g_tgt),
        _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_int(C), R(voxel_shift), R(eps),
        ctypes.c_int(bool(stop_grad)), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def trilinear_bwd_mask(vol, mask, src, tgt, raylen, gout, n_points=500, alphamin=None, alphamax=None, voxel_shift=0.5,
                       eps=1e-8, align_corners=False, want_vol=True, dtype=np.float64):
    """Backward of trilinear_fwd_mask for gout (B,C,N); same outputs as trilinear_bwd (fixed range + its partials)."""
    vol, mask, src, tgt, raylen, gout = _prep(dtype, vol, mask, src, tgt, raylen, gout)
    B, N, C = tgt.shape[0], tgt.shape[1], gout.shape[1]
    if alphamin is None or alphamax is None:
        alphamin, alphamax = alpha_minmax(vol.shape, src, tgt, voxel_shift, eps, dtype)
    g_src, g_tgt = np.zeros((B, 1, 3), dtype=dtype), np.zeros((B, N, 3), dtype=dtype)
    g_len = np.zeros((B, 1, N), dtype=dtype)
    g_vol = np.zeros(vol.shape, dtype=dtype) if want_vol else None
    R = _real(dtype)
    ga0, ga1 = R(0), R(0)
    getattr(lib(), "oracle_trilinear_bwd_mask_" + _suf(dtype))(
        _p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src), _p(g_tgt),
        _p(g_len), _p(g_vol), ctypes.byref(ga0), ctypes.byref(ga1), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_int(C),
        R(voxel_shift), R(eps), ctypes.c_int(n_points), R(alphamin), R(alphamax), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol, g_alphamin=ga0.value,
                g_alphamax=ga1.value)


def siddon_bilinear(vol, src, tgt, raylen, gout=None, voxel_shift=0.5, eps=1e-8, stop_grad=False, align_corners=False,
                    dtype=np.float64):
    """Siddon with mode="bilinear": image (B,1,N) and, when gout is given, the gradients dict of siddon_bwd."""
    if gout is None:
        vol, src, tgt, raylen = _prep(dtype, vol, src, tgt, raylen)
    else:
        vol, src, tgt, raylen, gout = _prep(dtype, vol, src, tgt, raylen, gout)
    B, N = tgt.shape[0], tgt.shape[1]
    out = np.zeros((B, 1, N), dtype=dtype)
    grads = gout is not None
    g_src = np.zeros((B, 1, 3), dtype=dtype) if grads else None
    g_tgt = np.zeros((B, N, 3), dtype=dtype) if grads else None
    g_len = np.zeros((B, 1, N), dtype=dtype) if grads else None
    g_vol = np.zeros(vol.shape, dtype=dtype) if (grads and not stop_grad) else None
    R = _real(dtype)
    getattr(lib(), "oracle_siddon_bilinear_" + _suf(dtype))(
        _p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(out), _p(g_src), _p(g_tgt),
        _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N), R(voxel_shift), R(eps), ctypes.c_int(bool(stop_grad)),
        ctypes.c_int(bool(align_corners)))
    return dict(img=out, g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)
