"""Synthetic CT volumes, geometry and poses used by the tests, the goldens and bench.py.

Restates the measurement set-up of SURVEY.md section 8(d): a 256 mm cube of D^3 isotropic voxels
with its isocentre at the world origin, AP reorientation (reference data.py:86-97), sdd = 1020 mm,
a 300 mm detector, and poses drawn like notebooks/tutorials/registration.ipynb cell 5.

Everything is generated on the host with numpy's PCG64 / torch's CPU generator, so the numbers are
identical in the build container (where the goldens are recorded) and on the GPU box.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CUBE_MM = 256.0
SDD_MM = 1020.0
DETECTOR_MM = 300.0

AP_REORIENT = [[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]


class _Image:
    def __init__(self, data: torch.Tensor, affine: np.ndarray):
        self.data = data
        self.affine = affine


class Subject:
    """Attribute bag with the fields reference drr.py:64-91 reads from a torchio.Subject."""

    def __init__(self, volume: torch.Tensor, affine: np.ndarray, reorient: torch.Tensor, mask=None):
        if volume.dim() == 3:
            volume = volume.unsqueeze(0)
        self.volume = _Image(volume, affine)
        self.density = _Image(volume, affine)
        self.mask = None if mask is None else _Image(mask if mask.dim() == 4 else mask.unsqueeze(0), affine)
        self.reorient = reorient
        self.orientation = "AP"
        self.fiducials = None


def make_volume(shape, kind: str = "rand", seed: int = 0) -> np.ndarray:
    """fp32 volume of `shape` (int or 3-tuple); 'rand' = U[0,1) noise, 'phantom' = blobs + hard sphere + 1% noise,
    'smooth' = the two Gaussian blobs only (for gradient checks: no edges a grazing ray could flip across)."""
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "rand":
        return rng.random(shape, dtype=np.float32)
    if kind not in ("phantom", "smooth"):
        raise ValueError(kind)
    ax = [np.linspace(-1.0, 1.0, n, dtype=np.float64) for n in shape]
    x, y, z = np.meshgrid(*ax, indexing="ij")
    vol = np.exp(-((x - 0.2) ** 2 + (y + 0.1) ** 2 + (z - 0.05) ** 2) / 0.18)
    vol += 0.6 * np.exp(-((x + 0.35) ** 2 + (y - 0.3) ** 2 + (z + 0.25) ** 2) / 0.05)
    if kind == "phantom":
        vol += 0.5 * ((x + 0.1) ** 2 + (y + 0.2) ** 2 + (z - 0.3) ** 2 < 0.09)
        vol += 0.01 * rng.random(shape)
    return vol.astype(np.float32)


def make_affine(shape, cube_mm: float = CUBE_MM) -> np.ndarray:
    """Voxel-index -> world (mm) affine with the isocentre at the origin (reference data.py:187-202)."""
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    sp = cube_mm / max(shape)
    aff = np.eye(4, dtype=np.float64)
    for a in range(3):
        aff[a, a] = sp
        aff[a, 3] = -sp * (shape[a] - 1) / 2.0
    return aff


def make_subject(volume, cube_mm: float = CUBE_MM, mask=None) -> Subject:
    vol = torch.as_tensor(volume)
    aff = make_affine(tuple(vol.shape[-3:]), cube_mm)
    return Subject(vol, aff, torch.tensor(AP_REORIENT, dtype=torch.float32), mask=mask)


def detector_kwargs(height: int, width: int | None = None) -> dict:
    width = height if width is None else width
    return dict(sdd=SDD_MM, height=height, width=width, delx=DETECTOR_MM / width, dely=DETECTOR_MM / height)


def make_poses(batch: int, seed: int = 0, rot_range: float = math.pi / 4, xyz_range: float = 30.0):
    """Euler-ZXY rotations (rad) and translations (mm): canonical AP view for batch==1, seeded jitter otherwise."""
    base = torch.tensor([[0.0, 850.0, 0.0]])
    if batch == 1:
        return torch.zeros(1, 3), base.clone()
    g = torch.Generator().manual_seed(seed)
    rot = (torch.rand(batch, 3, generator=g) * 2 - 1) * rot_range
    xyz = base + (torch.rand(batch, 3, generator=g) * 2 - 1) * xyz_range
    return rot, xyz
