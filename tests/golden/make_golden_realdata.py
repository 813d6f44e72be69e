#!/usr/bin/env python
"""Real-data golden: a crop of the reference's example CT (diffdrr/data/cxr.nii.gz, 512x512x133 int16 HU,
0.703 x 0.703 x 2.5 mm) rendered by the UNMODIFIED reference renderers after the reference's own HU -> density map.
Stored: the int16 crop, its density (reference transform_hu_to_density), rays of two poses on a 40^2 detector, fp32 / fp64
Siddon and trilinear images.  Run in the build container: python tests/golden/make_golden_realdata.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "_refshim"))
sys.path.insert(0, "/root/reference")
from diffdrr.renderers import Siddon, Trilinear  # noqa: E402  (reference)

# the reference's transform lives in data.py, which imports torchio for OTHER functions: execute only that function's source
src = open("/root/reference/diffdrr/data.py").read()
start = src.index("def transform_hu_to_density")
ns = {"torch": torch}
exec(compile(src[start:], "reference_data_py_tail", "exec"), ns)   # the unmodified function body
ref_transform = ns["transform_hu_to_density"]

from diffdrr_b200 import DRR, synthetic  # noqa: E402  (host geometry only: verified identical to the reference's)
from diffdrr_b200.data import read_nifti  # noqa: E402
from diffdrr_b200.pose import convert  # noqa: E402

vol, affine = read_nifti("/root/reference/diffdrr/data/cxr.nii.gz")
assert vol.shape == (512, 512, 133)
crop = np.ascontiguousarray(vol[160:256, 200:296, 40:104]).astype(np.int16)      # 96 x 96 x 64 voxels through the thorax
density = ref_transform(torch.from_numpy(crop.astype(np.float32)), 2.5).numpy()
sp = np.abs(np.diag(affine)[:3])
aff = np.diag([sp[0], sp[1], sp[2], 1.0])
aff[:3, 3] = -sp * (np.array(crop.shape) - 1) / 2.0                               # isocentre at the origin (data.py:187-202)
subj = synthetic.Subject(torch.from_numpy(density), aff, torch.tensor(synthetic.AP_REORIENT, dtype=torch.float32))
H = 40
drr = DRR(subj, sdd=1020.0, height=H, width=H, delx=4.0, dely=4.0)
rot = torch.tensor([[0.0, 0.0, 0.0], [0.35, -0.2, 0.15]])
xyz = torch.tensor([[0.0, 850.0, 0.0], [10.0, 800.0, -15.0]])
with torch.no_grad():
    s, t = drr.detector(convert(rot, xyz, parameterization="euler_angles", convention="ZXY"), None)
    raylen = (t - s).norm(dim=-1).unsqueeze(1)
    s, t = drr.affine_inverse(s), drr.affine_inverse(t)
rec = {"hu_crop": crop, "density_sub3": density[::3, ::3, ::3].copy(), "density_sum": np.float64(density.astype(np.float64).sum()),
       "affine": aff, "source": s.numpy(), "target": t.numpy(), "raylen": raylen.numpy(),
       "bone_attenuation_multiplier": np.float32(2.5)}
for dt, suf in ((torch.float32, "f32"), (torch.float64, "f64")):
    v = torch.from_numpy(density).to(dt)
    with torch.no_grad():
        rec[f"siddon_{suf}"] = Siddon()(v, s.to(dt), t.to(dt), raylen.to(dt)).numpy()
        rec[f"trilinear_{suf}"] = Trilinear()(v, s.to(dt), t.to(dt), raylen.to(dt), n_points=300).numpy()
np.savez_compressed(os.path.join(HERE, "realdata_cxr_crop.npz"), **rec)
print({k: getattr(v, "shape", v) for k, v in rec.items()}, os.path.getsize(os.path.join(HERE, "realdata_cxr_crop.npz")))
