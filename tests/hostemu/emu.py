"""ctypes front-end of tests/hostemu/libhostemu.so (product per-ray math compiled for the CPU; test infra)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostemu.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "hostemu.cpp")] + [
            os.path.join(_ROOT, "diffdrr_b200", "csrc", f) for f in ("ray_math.cuh", "common.cuh", "psync.cuh", "brick.cuh", "ncc_math.cuh")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                                   "-Wno-unknown-pragmas", "-o", _SO, srcs[0]])
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _common(vol, src, tgt, raylen):
    vol, src, tgt, raylen = _f(vol), _f(src), _f(tgt), _f(raylen)
    return vol, src, tgt, raylen, tgt.shape[0], tgt.shape[1]


def siddon_fwd(vol, src, tgt, raylen, voxel_shift=0.5, eps=1e-8, reduce="sum", align_corners=False, general=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    out = np.empty((B, 1, N), np.float32)
    fn = lib().emu_siddon_general if general else lib().emu_siddon_fwd
    fn(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B), ctypes.c_long(N),
       ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int({"sum": 0, "max": 1}[reduce]),
       ctypes.c_int(bool(align_corners)))
    return out


def siddon_fwd_ilp(vol, src, tgt, raylen, voxel_shift=0.5, eps=1e-8, unroll=4):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    out = np.empty((B, 1, N), np.float32)
    lib().emu_siddon_fwd_ilp(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                             ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                             ctypes.c_int(unroll))
    return out


def siddon_fwd_psync(vol, src, tgt, raylen, voxel_shift=0.5, eps=1e-8, slab=0, unroll=2):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    out = np.empty((B, 1, N), np.float32)
    lib().emu_siddon_fwd_psync(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                               ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                               ctypes.c_int(slab), ctypes.c_int(unroll))
    return out


def siddon_visits(shape, src, tgt, voxel_shift=0.5, eps=1e-8):
    src, tgt = _f(src), _f(tgt)
    B, N = tgt.shape[0], tgt.shape[1]
    out = np.empty((B, N), np.int32)
    lib().emu_siddon_visits(*map(ctypes.c_int, shape), _p(src), _p(tgt), _p(out), ctypes.c_int(B), ctypes.c_long(N),
                            ctypes.c_float(voxel_shift), ctypes.c_float(eps))
    return out


def siddon_bwd(vol, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False, lean_slab=None):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    args = (_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src), _p(g_tgt),
            _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
            ctypes.c_int(bool(stop_grad)))
    if lean_slab is None:
        lib().emu_siddon_bwd(*args)
    else:
        lib().emu_siddon_bwd_lean(*args, ctypes.c_int(lean_slab))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def siddon_sens(vol, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False, slab=0):
    """Fused forward + sensitivities + elementwise backward (the training-step fast path)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, out = np.zeros((B, 1, N), np.float32), np.zeros((B, 1, N), np.float32)
    lib().emu_siddon_sens(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src),
                          _p(g_tgt), _p(g_len), _p(out), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift),
                          ctypes.c_float(eps), ctypes.c_int(bool(stop_grad)), ctypes.c_int(slab))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, img=out)


def trilinear_fwd(vol, src, tgt, raylen, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8, reduce="sum",
                  align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    out = np.empty((B, 1, N), np.float32)
    lib().emu_trilinear_fwd(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                            ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                            ctypes.c_int(n_points), ctypes.c_float(alphamin), ctypes.c_float(alphamax),
                            ctypes.c_int({"sum": 0, "max": 1}[reduce]), ctypes.c_int(bool(align_corners)))
    return out


def trilinear_bwd(vol, src, tgt, raylen, gout, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8,
                  align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    g_ar = np.zeros(2, np.float32)
    lib().emu_trilinear_bwd(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src),
                            _p(g_tgt), _p(g_len), _p(g_vol), _p(g_ar), ctypes.c_int(B), ctypes.c_long(N),
                            ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(n_points),
                            ctypes.c_float(alphamin), ctypes.c_float(alphamax), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol, g_alphamin=float(g_ar[0]),
                g_alphamax=float(g_ar[1]))


def siddon_fwd_mask(vol, mask, src, tgt, raylen, C, voxel_shift=0.5, eps=1e-8):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    mask = _f(mask)
    out = np.empty((B, C, N), np.float32)
    lib().emu_siddon_fwd_mask(_p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                              ctypes.c_int(B), ctypes.c_long(N), ctypes.c_int(C), ctypes.c_float(voxel_shift), ctypes.c_float(eps))
    return out


def trilinear_fwd_mask(vol, mask, src, tgt, raylen, C, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8,
                       align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    mask = _f(mask)
    out = np.empty((B, C, N), np.float32)
    lib().emu_trilinear_fwd_mask(_p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                                 ctypes.c_int(B), ctypes.c_long(N), ctypes.c_int(C), ctypes.c_float(voxel_shift),
                                 ctypes.c_float(eps), ctypes.c_int(n_points), ctypes.c_float(alphamin),
                                 ctypes.c_float(alphamax), ctypes.c_int(bool(align_corners)))
    return out


def siddon_bwd_general(vol, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False, reduce="sum",
                       align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    lib().emu_siddon_bwd_general(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src),
                                 _p(g_tgt), _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N),
                                 ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(bool(stop_grad)),
                                 ctypes.c_int({"sum": 0, "max": 1}[reduce]), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def trilinear_bwd_max(vol, src, tgt, raylen, gout, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8,
                      align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    g_ar = np.zeros(2, np.float32)
    lib().emu_trilinear_bwd_max(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_src),
                                _p(g_tgt), _p(g_len), _p(g_vol), _p(g_ar), ctypes.c_int(B), ctypes.c_long(N),
                                ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(n_points),
                                ctypes.c_float(alphamin), ctypes.c_float(alphamax), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol, g_alphamin=float(g_ar[0]),
                g_alphamax=float(g_ar[1]))


def transposed_padded(vol, axis, pad=4):
    """Copy of `vol` with `axis` fastest (the other two axes keep their order), flattened and padded by `pad` floats."""
    order = [a for a in range(3) if a != axis] + [axis]
    flat = np.ascontiguousarray(np.transpose(_f(vol), order)).ravel()
    return np.concatenate([flat, np.zeros(pad, np.float32)])


def siddon_fwd_chunk(vol, src, tgt, raylen, axis, width=4, voxel_shift=0.5, eps=1e-8, slab=0):
    """EXPERIMENT: chunk-reuse lean walk; returns (chunk result, plain lean walk with the same slab cuts)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    volT = transposed_padded(vol, axis)
    out, ref = np.empty((B, 1, N), np.float32), np.empty((B, 1, N), np.float32)
    lib().emu_siddon_fwd_chunk(_p(volT), *map(ctypes.c_int, vol.shape), ctypes.c_int(axis), ctypes.c_int(width), _p(src),
                               _p(tgt), _p(raylen), _p(out), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift),
                               ctypes.c_float(eps), ctypes.c_int(slab))
    lib().emu_siddon_fwd_lean_slab(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(ref),
                                   ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                                   ctypes.c_int(slab))
    return out, ref


def siddon_sens_chunk(vol, src, tgt, raylen, axis, width=4, voxel_shift=0.5, eps=1e-8, slab=0):
    """EXPERIMENT: (out, sens) of the sensitivities walk through the chunk loader and through the plain loader."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    res = []
    for chunked, data in ((1, transposed_padded(vol, axis)), (0, np.concatenate([vol.ravel(), np.zeros(4, np.float32)]))):
        out, sens = np.empty((B, 1, N), np.float32), np.empty((B, N, 8), np.float32)
        lib().emu_siddon_sens_chunk(_p(data), *map(ctypes.c_int, vol.shape), ctypes.c_int(axis), ctypes.c_int(width), _p(src),
                                    _p(tgt), _p(raylen), _p(out), _p(sens), ctypes.c_int(B), ctypes.c_long(N),
                                    ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(slab), ctypes.c_int(chunked))
        res.append((out, sens))
    return res


def siddon_bilinear(vol, src, tgt, raylen, gout=None, voxel_shift=0.5, eps=1e-8, stop_grad=False, reduce="sum",
                    align_corners=False):
    """Siddon(mode="bilinear"): dict(img, and with gout the gradients)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    grads = gout is not None
    gout = _f(gout) if grads else np.zeros((B, 1, N), np.float32)
    out = np.zeros((B, 1, N), np.float32)
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    lib().emu_siddon_bilinear(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(out),
                              _p(g_src), _p(g_tgt), _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N),
                              ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(bool(stop_grad)),
                              ctypes.c_int({"sum": 0, "max": 1}[reduce]), ctypes.c_int(bool(align_corners)),
                              ctypes.c_int(bool(grads)))
    return dict(img=out, g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def siddon_bwd_mask(vol, mask, src, tgt, raylen, gout, voxel_shift=0.5, eps=1e-8, stop_grad=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    mask, gout = _f(mask), _f(gout)
    C = gout.shape[1]
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    lib().emu_siddon_bwd_mask(_p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout),
                              _p(g_src), _p(g_tgt), _p(g_len), _p(g_vol), ctypes.c_int(B), ctypes.c_long(N), ctypes.c_int(C),
                              ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(bool(stop_grad)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol)


def trilinear_bwd_mask(vol, mask, src, tgt, raylen, gout, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8,
                       align_corners=False):
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    mask, gout = _f(mask), _f(gout)
    C = gout.shape[1]
    g_src, g_tgt = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32)
    g_len, g_vol = np.zeros((B, 1, N), np.float32), np.zeros(vol.shape, np.float32)
    g_ar = np.zeros(2, np.float32)
    lib().emu_trilinear_bwd_mask(_p(vol), _p(mask), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout),
                                 _p(g_src), _p(g_tgt), _p(g_len), _p(g_vol), _p(g_ar), ctypes.c_int(B), ctypes.c_long(N),
                                 ctypes.c_int(C), ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(n_points),
                                 ctypes.c_float(alphamin), ctypes.c_float(alphamax), ctypes.c_int(bool(align_corners)))
    return dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_volume=g_vol, g_alphamin=float(g_ar[0]),
                g_alphamax=float(g_ar[1]))


def trilinear_packed(vol, src, tgt, raylen, gout, n_points, alphamin, alphamax, voxel_shift=0.5, eps=1e-8, slab=0):
    """Forward + pose-gradient backward through the packed-corner volume; returns (out, grads dict)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    gout = _f(gout)
    out = np.empty((B, 1, N), np.float32)
    g_src, g_tgt, g_len, g_ar = np.zeros((B, 1, 3), np.float32), np.zeros((B, N, 3), np.float32), np.zeros((B, 1, N), np.float32), np.zeros(2, np.float32)
    lib().emu_trilinear_packed(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(out),
                               _p(g_src), _p(g_tgt), _p(g_len), _p(g_ar), ctypes.c_int(B), ctypes.c_long(N),
                               ctypes.c_float(voxel_shift), ctypes.c_float(eps), ctypes.c_int(n_points),
                               ctypes.c_float(alphamin), ctypes.c_float(alphamax), ctypes.c_int(slab))
    return out, dict(g_source=g_src, g_target=g_tgt, g_raylen=g_len, g_alphamin=float(g_ar[0]), g_alphamax=float(g_ar[1]))


def siddon_fwd_brick(vol, src, tgt, raylen, H, W, brick=(24, 32, 32), voxel_shift=0.5, eps=1e-8, check=False):
    """Brick-major forward exactly as siddon_brick.cu decomposes it (full H x W grid).  Returns (image, violations,
    stats) -- violations counts hits the detector rectangle / conservative test would have dropped (must be 0)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    assert N == H * W
    out = np.empty((B, 1, N), np.float32)
    stats = np.zeros(4, np.int64)
    fn = lib().emu_siddon_fwd_brick
    fn.restype = ctypes.c_long
    viol = fn(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
              ctypes.c_int(H), ctypes.c_int(W), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
              *map(ctypes.c_int, brick), ctypes.c_int(int(check)), _p(stats))
    return out, int(viol), dict(zip(("candidates", "maybe", "exact", "walked"), stats.tolist()))


def siddon_fwd_brick2(vol, src, tgt, raylen, H, W, brick=(24, 32, 32), voxel_shift=0.5, eps=1e-8, check=False, lean=True):
    """Production brick decomposition (outline-clipped tile bands; lean=True: set-up without fix-ups + accumulated alphas)."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    assert N == H * W
    out = np.empty((B, 1, N), np.float32)
    stats = np.zeros(4, np.int64)
    fn = lib().emu_siddon_fwd_brick2
    fn.restype = ctypes.c_long
    viol = fn(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out), ctypes.c_int(B),
              ctypes.c_int(H), ctypes.c_int(W), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
              *map(ctypes.c_int, brick), ctypes.c_int(int(check)), ctypes.c_int(int(lean)), _p(stats))
    return out, int(viol), dict(zip(("candidates", "maybe", "exact", "walked"), stats.tolist()))


def siddon_bwd_vol_brick(vol_shape, src, tgt, raylen, gout, H, W, brick=(24, 32, 32), voxel_shift=0.5, eps=1e-8):
    """Volume gradient through the brick decomposition (siddon_brick.cu, BWD mode); g_vol (D0, D1, D2) overwritten."""
    src, tgt, raylen, gout = _f(src), _f(tgt), _f(raylen), _f(gout)
    B, N = tgt.shape[0], tgt.shape[1]
    assert N == H * W
    g_vol = np.full(tuple(vol_shape), np.nan, np.float32)
    lib().emu_siddon_bwd_vol_brick(*map(ctypes.c_int, vol_shape), _p(src), _p(tgt), _p(raylen), _p(gout), _p(g_vol), ctypes.c_int(B),
                                   ctypes.c_int(H), ctypes.c_int(W), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                                   *map(ctypes.c_int, brick))
    return g_vol


def ncc(x1, x2, gscore=None, eps=1e-5):
    """ncc.cu's math (chunked double moments, closed-form gradient) on the CPU: (score [B], g_x1, g_x2) for x [B, C, H, W]."""
    x1, x2 = _f(x1), _f(x2)
    B, C = x1.shape[:2]
    N = int(np.prod(x1.shape[2:]))
    stats = np.empty((B * C, 8), np.float32)
    score = np.empty(B, np.float32)
    lib().emu_ncc_fwd(_p(x1), _p(x2), ctypes.c_int(B), ctypes.c_int(C), ctypes.c_long(N), ctypes.c_float(eps), _p(stats), _p(score))
    if gscore is None:
        return score
    g1, g2 = np.empty_like(x1), np.empty_like(x2)
    lib().emu_ncc_bwd(_p(x1), _p(x2), _p(stats), _p(_f(gscore)), _p(g1), _p(g2), ctypes.c_int(B), ctypes.c_int(C), ctypes.c_long(N))
    return score, g1, g2


def siddon_fwd_lean_pieces(vol, src, tgt, raylen, pieces, voxel_shift=0.5, eps=1e-8):
    """Lean forward walk cut into `pieces` along each ray's own major axis (the MAJ small-batch kernels); pieces=0: uncut."""
    vol, src, tgt, raylen, B, N = _common(vol, src, tgt, raylen)
    out = np.empty((B, 1, N), np.float32)
    lib().emu_siddon_fwd_lean_slab(_p(vol), *map(ctypes.c_int, vol.shape), _p(src), _p(tgt), _p(raylen), _p(out),
                                   ctypes.c_int(B), ctypes.c_long(N), ctypes.c_float(voxel_shift), ctypes.c_float(eps),
                                   ctypes.c_int(-int(pieces)))
    return out
