"""Image similarity used by the 2D/3D registration loop (the path's caller).

Only normalized cross correlation is provided -- the loss of the reference's registration tutorial
(reference diffdrr/metrics.py:21-44, incl. its patch mode, metrics.py:16-19,30).  The rest of the reference's loss zoo
(multiscale / gradient NCC, mutual information, geodesics) is outside the projector path.
"""
from __future__ import annotations

import torch


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
        return (self.norm(x1) * self.norm(x2)).flatten(1).sum(dim=1) / (c * h * w)
