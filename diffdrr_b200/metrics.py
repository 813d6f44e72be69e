"""Image similarity used by the 2D/3D registration loop (the path's caller).

Only normalized cross correlation is provided -- the loss of the reference's registration tutorial
(reference diffdrr/metrics.py:21-44, incl. its patch mode, metrics.py:16-19,30).  The rest of the reference's loss zoo
(multiscale / gradient NCC, mutual information, geodesics) is outside the projector path.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib


class _NCCFunction(torch.autograd.Function):
    """score[b] = mean_{c,h,w} norm(x1) norm(x2) through b200drr_ncc_fwd / _bwd (three launches per training step instead of
    the ~40 of the reference's elementwise graph, metrics.py:21-44): the registration loop at one pose per step is launch-bound."""

    @staticmethod
    def forward(ctx, x1, x2, eps):
        x1, x2 = x1.contiguous(), x2.contiguous()
        B, C = x1.shape[0], x1.shape[1]
        N = x1[0, 0].numel()
        lib = _lib.load()
        ws = torch.empty(int(lib.b200drr_ncc_workspace_bytes(B, C, N)) // 8, dtype=torch.float64, device=x1.device)
        stats = torch.empty(B * C, 8, dtype=torch.float32, device=x1.device)
        score = torch.empty(B, dtype=torch.float32, device=x1.device)
        _lib.check(lib.b200drr_ncc_fwd(x1.data_ptr(), x2.data_ptr(), B, C, N, float(eps), ws.data_ptr(), stats.data_ptr(),
                                       score.data_ptr(), torch.cuda.current_stream(x1.device).cuda_stream), "b200drr_ncc_fwd")
        ctx.save_for_backward(x1, x2, stats)
        return score

    @staticmethod
    @once_differentiable
    def backward(ctx, gscore):
        x1, x2, stats = ctx.saved_tensors
        B, C = x1.shape[0], x1.shape[1]
        N = x1[0, 0].numel()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need1 or need2):
            return None, None, None
        g1 = torch.empty_like(x1) if need1 else None
        g2 = torch.empty_like(x2) if need2 else None
        gscore = gscore.contiguous().float()
        _lib.check(_lib.load().b200drr_ncc_bwd(x1.data_ptr(), x2.data_ptr(), stats.data_ptr(), gscore.data_ptr(),
                                               g1.data_ptr() if need1 else None, g2.data_ptr() if need2 else None, B, C, N,
                                               torch.cuda.current_stream(x1.device).cuda_stream), "b200drr_ncc_bwd")
        return g1, g2, None


class NormalizedCrossCorrelation2d(torch.nn.Module):
    """Zero-normalized cross correlation between two image batches (B, C, H, W) -> (B,)."""

    def __init__(self, patch_size=None, eps: float = 1e-5):
        super().__init__()
        self.patch_size = patch_size
        self.eps = eps

    @staticmethod
    def to_patches(x, patch_size: int):
        """(B, C, H, W) -> (B, C * H' * W', p, p): every p x p window (stride 1) becomes a channel (metrics.py:16-19)."""
        x = x.unfold(2, patch_size, 1).unfold(3, patch_size, 1)
        return x.reshape(x.shape[0], -1, patch_size, patch_size)

    def fused_ok(self, x1, x2) -> bool:
        """The CUDA kernels take full images (patch_size None: the patch view is a strided unfold) of fp32 on one device."""
        return (self.patch_size is None and x1.is_cuda and x2.is_cuda and x1.device == x2.device and x1.dtype == torch.float32
                and x2.dtype == torch.float32 and x1.dim() == 4 and x1.shape[0] * x1.shape[1] <= 65535 and x1.numel() > 0)

    def norm(self, x):
        mu = x.mean(dim=(-1, -2), keepdim=True)
        var = x.var(dim=(-1, -2), keepdim=True, correction=0) + self.eps
        return (x - mu) / var.sqrt()

    def forward(self, x1, x2):
        if self.patch_size is not None:
            x1, x2 = self.to_patches(x1, self.patch_size), self.to_patches(x2, self.patch_size)
        if x1.shape != x2.shape:
            raise AssertionError("Input images must be the same size")
        _, c, h, w = x1.shape
        if self.fused_ok(x1, x2):
            with torch.cuda.device(x1.device):
                return _NCCFunction.apply(x1, x2, self.eps)
        return (self.norm(x1) * self.norm(x2)).flatten(1).sum(dim=1) / (c * h * w)
