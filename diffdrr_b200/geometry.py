"""Per-pose algebra of the fused pose-in path as two CUDA kernels each way (include/b200drr.h: b200drr_euler_pose_*,
b200drr_pose_rays_*), instead of ~95 tiny ATen kernels per training step.

    pose parameters --euler_pose--> P (B,4,4) --pose_rays--> (src, G, Wd) --> b200drr_siddon_fwd_sens_pose / _fwd_pose

Restates `convert(..., parameterization="euler_angles")` (reference pose.py) and `Detector.forward` + the ray-length /
affine_inverse lines of `DRR.render` (detector.py:144-154, drr.py:201-205).  Only used on CUDA fp32 tensors by
`DRR.forward`; the public `convert` / `Detector` stay plain torch.
"""
from __future__ import annotations

import math

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .renderers import _ptr, _stream

_AXIS = {"X": 0, "Y": 1, "Z": 2}


def euler_convention_ok(convention) -> bool:
    return (isinstance(convention, str) and len(convention) == 3 and all(c in _AXIS for c in convention)
            and convention[1] not in (convention[0], convention[2]))


class _EulerPoseFunction(torch.autograd.Function):
    """(rot (B,3), xyz (B,3)) -> pose matrix (B,4,4), R = R_c0 R_c1 R_c2, translation column R.xyz."""

    @staticmethod
    def forward(ctx, rot, xyz, axes, scale):
        rot, xyz = rot.contiguous().float(), xyz.contiguous().float()
        B = rot.shape[0]
        if rot.shape != xyz.shape or rot.dim() != 2 or rot.shape[1] != 3:
            # the C ABI only receives B: a (1, 3) translation next to (B, 3) rotations would be read out of bounds
            raise ValueError(f"euler_pose needs rot and xyz of the same (B, 3) shape, got {tuple(rot.shape)} and {tuple(xyz.shape)}")
        P = torch.empty(B, 4, 4, dtype=torch.float32, device=rot.device)
        with torch.cuda.device(rot.device):
            _lib.check(_lib.load().b200drr_euler_pose_fwd(_ptr(rot), _ptr(xyz), *axes, scale, _ptr(P), B, _stream()),
                       "b200drr_euler_pose_fwd")
        ctx.save_for_backward(rot, xyz)
        ctx.cfg = (axes, scale)
        return P

    @staticmethod
    @once_differentiable
    def backward(ctx, gP):
        rot, xyz = ctx.saved_tensors
        axes, scale = ctx.cfg
        B = rot.shape[0]
        gP = gP.contiguous().float()
        g_rot = torch.empty_like(rot) if ctx.needs_input_grad[0] else None
        g_xyz = torch.empty_like(xyz) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(rot.device):
            _lib.check(_lib.load().b200drr_euler_pose_bwd(_ptr(rot), _ptr(xyz), *axes, scale, _ptr(gP), _ptr(g_rot),
                                                          _ptr(g_xyz), B, _stream()), "b200drr_euler_pose_bwd")
        return g_rot, g_xyz, None, None


def euler_pose(rot: torch.Tensor, xyz: torch.Tensor, convention: str, degrees: bool = False) -> torch.Tensor:
    axes = tuple(_AXIS[c] for c in convention)
    return _EulerPoseFunction.apply(rot, xyz, axes, math.pi / 180.0 if degrees else 1.0)


class _PoseRaysFunction(torch.autograd.Function):
    """P (B,4,4) -> src (B,3), G (B,3,4), Wd (B,3,4) for constant Q = reorient.calibration, r = reorient[:,3], Ainv."""

    @staticmethod
    def forward(ctx, P, Q, r, Ainv):
        P = P.contiguous().float()
        B = P.shape[0]
        dev = P.device
        src = torch.empty(B, 3, dtype=torch.float32, device=dev)
        G = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
        Wd = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200drr_pose_rays_fwd(_ptr(P), _ptr(Q), _ptr(r), _ptr(Ainv), _ptr(src), _ptr(G), _ptr(Wd), B,
                                                         _stream()), "b200drr_pose_rays_fwd")
        ctx.save_for_backward(Q, r, Ainv)
        return src, G, Wd

    @staticmethod
    @once_differentiable
    def backward(ctx, g_src, g_G, g_Wd):
        Q, r, Ainv = ctx.saved_tensors
        B = g_G.shape[0]
        g_src, g_G, g_Wd = g_src.contiguous().float(), g_G.contiguous().float(), g_Wd.contiguous().float()
        gP = torch.empty(B, 4, 4, dtype=torch.float32, device=g_G.device)
        with torch.cuda.device(g_G.device):
            _lib.check(_lib.load().b200drr_pose_rays_bwd(_ptr(Q), _ptr(r), _ptr(Ainv), _ptr(g_src), _ptr(g_G), _ptr(g_Wd),
                                                         _ptr(gP), B, _stream()), "b200drr_pose_rays_bwd")
        return gP, None, None, None


def pose_rays(P: torch.Tensor, Q: torch.Tensor, r: torch.Tensor, Ainv: torch.Tensor):
    return _PoseRaysFunction.apply(P, Q, r, Ainv)
