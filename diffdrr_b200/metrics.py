"""Image similarity used by the 2D/3D registration loop (the path's caller).

Only normalized cross correlation is provided -- the loss of the reference's registration tutorial
(reference diffdrr/metrics.py:21-44, without the optional patch mode).  The rest of the reference's loss zoo
(multiscale / gradient NCC, mutual information, geodesics) is outside the projector path.
"""
from __future__ import annotations

import torch


class NormalizedCrossCorrelation2d(torch.nn.Module):
    """Zero-normalized cross correlation between two image batches (B, C, H, W) -> (B,)."""

    def __init__(self, patch_size=None, eps: float = 1e-5):
        super().__init__()
        if patch_size is not None:
            raise NotImplementedError("patch-wise NCC is not provided by diffdrr_b200")
        self.patch_size = patch_size
        self.eps = eps

    def norm(self, x):
        mu = x.mean(dim=(-1, -2), keepdim=True)
        var = x.var(dim=(-1, -2), keepdim=True, correction=0) + self.eps
        return (x - mu) / var.sqrt()

    def forward(self, x1, x2):
        if x1.shape != x2.shape:
            raise AssertionError("Input images must be the same size")
        _, c, h, w = x1.shape
        return (self.norm(x1) * self.norm(x2)).flatten(1).sum(dim=1) / (c * h * w)
