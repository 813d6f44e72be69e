"""`Registration`: pose parameters as nn.Parameters on top of a DRR module (reference diffdrr/registration.py:14-50).

`PoseRegressor` (a timm CNN) is not part of the projector path and is not provided.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .pose import convert


class Registration(nn.Module):
    """Automatic 2D-to-3D registration by differentiable rendering: holds the pose being optimised."""

    def __init__(self, drr, rotation: torch.Tensor, translation: torch.Tensor, parameterization: str,
                 convention: str | None = None):
        super().__init__()
        self.drr = drr
        self._rotation = nn.Parameter(rotation)
        self._translation = nn.Parameter(translation)
        self.parameterization = parameterization
        self.convention = convention

    def forward(self, **kwargs):
        # same pose as `self.drr(self.pose)` (registration.py:34-35); handing DRR.forward the parameters lets it run the
        # parameter -> pose-matrix algebra as one kernel on its fused path
        return self.drr(self._rotation, self._translation, parameterization=self.parameterization,
                        convention=self.convention, **kwargs)

    @property
    def pose(self):
        return convert(self._rotation, self._translation, parameterization=self.parameterization,
                       convention=self.convention)

    @property
    def rotation(self):
        return self._rotation

    @property
    def translation(self):
        return self._translation
