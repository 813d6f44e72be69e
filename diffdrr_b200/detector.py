"""C-arm geometry: turns an SE(3) pose into X-ray source / detector-pixel points in world coordinates.

Mirrors reference diffdrr/detector.py:17-202 (constructor arguments, buffers `source`, `target`, `_reorient`,
`_calibration`, the sign quirk of the `x0`/`y0` properties, `forward(extrinsic, calibration)`), restated --
the ray generation is O(B N) flops of differentiable PyTorch and stays on the host side of the boundary.
"""
from __future__ import annotations

import torch

from .pose import RigidTransform


class Detector(torch.nn.Module):
    """6-DoF X-ray source + flat-panel detector.  Pixel n = h * width + w (row-major, detector.py:126)."""

    def __init__(self, sdd: float, height: int, width: int, delx: float, dely: float, x0: float, y0: float,
                 reorient: torch.Tensor, n_subsample: int | None = None, reverse_x_axis: bool = False):
        super().__init__()
        self.height = height
        self.width = width
        self.n_subsample = n_subsample
        if n_subsample is not None:
            self.subsamples = []
        self.reverse_x_axis = reverse_x_axis
        source, target = self._initialize_carm()
        self.register_buffer("source", source)
        self.register_buffer("target", target)
        self.register_buffer("_reorient", reorient)
        calib = torch.eye(4)
        calib[0, 0], calib[1, 1], calib[2, 2] = delx, dely, sdd
        calib[0, 3], calib[1, 3] = x0, y0
        self.register_buffer("_calibration", calib)

    # intrinsic parameters live in the calibration matrix (detector.py:46-80); x0/y0 are reported NEGATED (quirk Q9)
    @property
    def sdd(self):
        return self._calibration[2, 2].item()

    @property
    def delx(self):
        return self._calibration[0, 0].item()

    @property
    def dely(self):
        return self._calibration[1, 1].item()

    @property
    def x0(self):
        return -self._calibration[0, -1].item()

    @property
    def y0(self):
        return -self._calibration[1, -1].item()

    @property
    def reorient(self):
        return RigidTransform(self._reorient)

    @property
    def calibration(self):
        return RigidTransform(self._calibration)

    @property
    def intrinsic(self):
        return make_intrinsic_matrix(self).to(self.source)

    def _initialize_carm(self):
        """Canonical C-arm: source at the origin, unit-spaced pixel grid on the plane z = 1 (detector.py:97-138)."""
        def centred(n):  # pixel centres, symmetric about 0 for even n, offset by half a pixel otherwise
            first = -((n + 1) // 2) + (1.0 if n % 2 else 0.5)  # == (-n // 2) + offset in Python floor division
            return -(torch.arange(n, dtype=torch.float32) + first)

        rows = centred(self.height)  # detector y, one value per image row h
        cols = centred(self.width)   # detector x, one value per image column w
        if not self.reverse_x_axis:
            cols = -cols
        yy, xx = torch.meshgrid(rows, cols, indexing="ij")
        target = torch.stack([xx, yy, torch.ones_like(xx)], dim=-1).reshape(1, -1, 3)
        source = torch.zeros(1, 1, 3)
        # canonical targets of the FULL grid's pixels (0,0), (0,W-1), (H-1,0): the brick-major kernel derives each pose's
        # detector plane from them when only a sub-sample of the pixels is rendered
        self.register_buffer("_corner_points", target[:, [0, self.width - 1, (self.height - 1) * self.width], :].clone(),
                             persistent=False)
        if self.n_subsample is not None:
            pick = torch.randperm(self.height * self.width)[: int(self.n_subsample)]
            target = target[:, pick, :]
            self.subsamples.append(pick.tolist())
        return source, target

    def corner_targets(self, extrinsic: RigidTransform, calibration: RigidTransform | None):
        """World-space targets (B, 3, 3) of the full grid's pixels (0,0), (0,W-1), (H-1,0) for the given poses."""
        calib = self.calibration if calibration is None else calibration
        pose = self.reorient.compose(extrinsic)
        return pose(calib(self._corner_points))

    def pick_tensor(self):
        """The current sub-sample's pixel indices as a cached int64 tensor on the detector's device (indexing with the Python
        list re-uploads it on every call: milliseconds per render)."""
        cached = getattr(self, "_pick_cache", None)
        if cached is None or cached[0] != len(self.subsamples) or cached[1].device != self.source.device:
            cached = (len(self.subsamples), torch.as_tensor(self.subsamples[-1], dtype=torch.int64, device=self.source.device))
            object.__setattr__(self, "_pick_cache", cached)
        return cached[1]

    def pixel_index(self):
        """(H*W,) int32 on the detector's device: position of every pixel in the current sub-sample, -1 if not sampled."""
        pick = self.pick_tensor()
        idx = torch.full((self.height * self.width,), -1, dtype=torch.int32, device=self.source.device)
        idx[pick] = torch.arange(len(pick), dtype=torch.int32, device=self.source.device)
        return idx

    def forward(self, extrinsic: RigidTransform, calibration: RigidTransform | None):
        """-> source (B,1,3), target (B,N,3) in world coordinates (detector.py:144-154)."""
        calib = self.calibration if calibration is None else calibration
        pose = self.reorient.compose(extrinsic)
        return pose(self.source), pose(calib(self.target))


def get_focal_length(intrinsic, delx: float, dely: float) -> float:
    return abs(intrinsic[0, 0] * delx + intrinsic[1, 1] * dely).item() / 2.0


def get_principal_point(intrinsic, height: int, width: int, delx: float, dely: float):
    x0 = delx * (intrinsic[0, 2] - width / 2)
    y0 = dely * (intrinsic[1, 2] - height / 2)
    return x0.item(), y0.item()


def parse_intrinsic_matrix(intrinsic, height: int, width: int, delx: float, dely: float):
    return (get_focal_length(intrinsic, delx, dely), *get_principal_point(intrinsic, height, width, delx, dely))


def make_intrinsic_matrix(detector: Detector):
    K = torch.eye(3)
    K[0, 0] = detector.sdd / detector.delx
    K[1, 1] = detector.sdd / detector.dely
    K[0, 2] = detector.x0 / detector.delx + detector.width / 2
    K[1, 2] = detector.y0 / detector.dely + detector.height / 2
    return K
