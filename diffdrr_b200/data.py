"""Minimal data path for real CTs: a NIfTI-1 reader (gzip + numpy, no nibabel / torchio) and the
Hounsfield-unit -> density map the reference applies before rendering (reference diffdrr/data.py:214-227).

Only what the projector path needs: the volume array, its voxel -> world affine, and the density transform.  The
reference's `read` (torchio Subject assembly, orientation handling, labelmaps, fiducials: data.py:44-211) is out of scope;
`synthetic.make_subject` builds the attribute bag `DRR` consumes from (density, affine).
"""
from __future__ import annotations

import gzip
import struct

import numpy as np
import torch

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32}


def read_nifti(path: str):
    """(array in file order (i fastest in the file -> returned as [i][j][k]), 4x4 affine) of a NIfTI-1 .nii / .nii.gz.
    Intensities are scaled with scl_slope / scl_inter when the header sets them; the affine is the sform (qform / pixdim
    fall-backs are not implemented: raises)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if struct.unpack("<i", raw[:4])[0] != 348:
        raise ValueError("not a little-endian NIfTI-1 file")
    dim = struct.unpack("<8h", raw[40:56])
    if dim[0] < 3:
        raise ValueError("need a 3-D volume")
    datatype = struct.unpack("<h", raw[70:72])[0]
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    slope, inter = struct.unpack("<2f", raw[112:120])
    sform_code = struct.unpack("<h", raw[254:256])[0]
    if datatype not in _NIFTI_DTYPES:
        raise NotImplementedError(f"NIfTI datatype {datatype}")
    if sform_code <= 0:
        raise NotImplementedError("only sform affines are read")
    affine = np.eye(4)
    affine[:3] = np.array(struct.unpack("<12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    n = dim[1] * dim[2] * dim[3]
    data = np.frombuffer(raw, dtype=_NIFTI_DTYPES[datatype], count=n, offset=vox_offset)
    vol = data.reshape(dim[3], dim[2], dim[1]).transpose(2, 1, 0)  # file order is i fastest: make it [i][j][k]
    if slope not in (0.0, 1.0) or inter != 0.0:
        vol = vol.astype(np.float32) * (slope if slope != 0.0 else 1.0) + inter
    return np.ascontiguousarray(vol), affine


def transform_hu_to_density(volume: torch.Tensor, bone_attenuation_multiplier: float = 1.0) -> torch.Tensor:
    """HU -> [0, 1] density (data.py:214-227): air (<= -800 HU) is clamped to the softest soft-tissue value, bone (> 350 HU)
    is scaled by `bone_attenuation_multiplier`, then the result is shifted / scaled to [0, 1]."""
    v = volume.to(torch.float32)
    soft = (v > -800) & (v <= 350)
    bone = v > 350
    floor = v[soft].min()
    density = torch.where(bone, v * bone_attenuation_multiplier, torch.where(soft, v, floor))
    density = density - density.min()
    return density / density.max()
