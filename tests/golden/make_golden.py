#!/usr/bin/env python
"""Record golden vectors by running the UNMODIFIED reference (DiffDRR @ /root/reference) on CPU.

The reference ships no numeric assertion on any rendered value, so parity is
pinned by running the reference itself.  /root/reference does not exist on the GPU box; this script
is run once in the build container and its output (`tests/golden/*.npz`) is committed:

    python tests/golden/make_golden.py

What is imported from the reference (all unmodified):
    diffdrr.renderers.Siddon / Trilinear  (renderers.py:11-254)   -- needs only torch
    diffdrr.drr.DRR, diffdrr.pose.convert (drr.py, detector.py, pose.py) -- through tests/_refshim
Every case stores the exact fp32 inputs fed to the reference and its fp32 AND fp64 outputs (the
fp64 run uses the same fp32 inputs promoted to double), plus autograd gradients of
loss = sum(w * img) for a stored weight image w.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from diffdrr.drr import DRR as RefDRR  # noqa: E402
from diffdrr.renderers import Siddon as RefSiddon, Trilinear as RefTrilinear  # noqa: E402
from torchio import Subject as ShimSubject  # noqa: E402

from diffdrr_b200 import synthetic  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def ref_subject(vol: np.ndarray):
    v = torch.from_numpy(vol)[None]
    return ShimSubject(v, synthetic.make_affine(vol.shape), torch.tensor(synthetic.AP_REORIENT))


def ref_rays(vol, height, width, rot, xyz, **kw):
    """source/target in voxel coordinates + ray lengths, exactly as reference drr.py:174,201-205 makes them."""
    drr = RefDRR(ref_subject(vol), **synthetic.detector_kwargs(height, width), **kw)
    from diffdrr.pose import convert

    pose = convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    source, target = drr.detector(pose, None)
    img = (target - source).norm(dim=-1).unsqueeze(1)
    source = drr.affine_inverse(source)
    target = drr.affine_inverse(target)
    return source.contiguous(), target.contiguous(), img.contiguous()


def run_renderer(renderer_cls, ctor_kw, fwd_kw, vol, source, target, raylen, w, dtype, grads=True):
    r = renderer_cls(**ctor_kw)
    v = torch.from_numpy(vol).to(dtype).requires_grad_(grads)
    s = source.to(dtype).clone().requires_grad_(grads)
    t = target.to(dtype).clone().requires_grad_(grads)
    l = raylen.to(dtype).clone().requires_grad_(grads)
    img = r(v, s, t, l, **fwd_kw)
    out = {"img": img.detach().numpy()}
    if grads:
        (img * w.to(dtype)).sum().backward()
        for name, x in (("g_volume", v), ("g_source", s), ("g_target", t), ("g_raylen", l)):
            out[name] = None if x.grad is None else x.grad.numpy()
    return out


def renderer_case(name, renderer_cls, vol, source, target, raylen, ctor_kw=None, fwd_kw=None, grads=True, store_vol="nc"):
    ctor_kw = ctor_kw or {}
    fwd_kw = fwd_kw or {}
    g = torch.Generator().manual_seed(1234)
    B, N = target.shape[0], target.shape[1]
    w = torch.rand(B, 1, N, generator=g, dtype=torch.float64)
    rec = {
        "source": source.numpy(),
        "target": target.numpy(),
        "raylen": raylen.numpy(),
        "w": w.numpy(),
    }
    rec["volume_key"] = np.str_(store_vol)
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        out = run_renderer(renderer_cls, ctor_kw, fwd_kw, vol, source, target, raylen, w, dtype, grads)
        for k, v in out.items():
            if v is not None:
                rec[f"{k}_{tag}"] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: img max {np.abs(rec['img_f64']).max():.4g}  "
          f"f32-vs-f64 {np.abs(rec['img_f32'] - rec['img_f64']).max() / np.abs(rec['img_f64']).max():.2e}  "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def drr_case(name, vol, height, rot, xyz, renderer, fwd_kw=None, store_vol="p32", **drr_kw):
    fwd_kw = fwd_kw or {}
    g = torch.Generator().manual_seed(4321)
    B = rot.shape[0]
    w = torch.rand(B, 1, height, height, generator=g, dtype=torch.float64)
    rec = {"rot": rot.numpy(), "xyz": xyz.numpy(), "w": w.numpy(), "height": np.int64(height)}
    rec["volume_key"] = np.str_(store_vol)
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        drr = RefDRR(ref_subject(vol), **synthetic.detector_kwargs(height), renderer=renderer, **drr_kw).to(dtype)
        r = rot.to(dtype).clone().requires_grad_(True)
        x = xyz.to(dtype).clone().requires_grad_(True)
        img = drr(r, x, parameterization="euler_angles", convention="ZXY", **fwd_kw)
        (img * w.to(dtype)).sum().backward()
        rec[f"img_{tag}"] = img.detach().numpy()
        rec[f"g_rot_{tag}"] = r.grad.numpy()
        rec[f"g_xyz_{tag}"] = x.grad.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    err = lambda a, b: np.abs(a - b).max() / np.abs(b).max()  # noqa: E731
    print(f"{name}: img f32-vs-f64 {err(rec['img_f32'], rec['img_f64']):.2e}  "
          f"g_rot {err(rec['g_rot_f32'], rec['g_rot_f64']):.2e}  g_xyz {err(rec['g_xyz_f32'], rec['g_xyz_f64']):.2e}  "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    # volumes are stored once (volumes.npz) and referenced by key; "rand64" is regenerated from its seed
    np.savez_compressed(os.path.join(HERE, "volumes.npz"),
                        nc=synthetic.make_volume((24, 32, 40), "phantom", seed=1),
                        p32=synthetic.make_volume(32, "phantom", seed=2))
    # ---- renderer-level cases on a NON-cubic volume (catches axis mix-ups) ------------------------------
    vol_nc = synthetic.make_volume((24, 32, 40), "phantom", seed=1)
    rot4, xyz4 = synthetic.make_poses(4, seed=0)
    s, t, l = ref_rays(vol_nc, 18, 18, rot4, xyz4)
    renderer_case("siddon_nc_b4", RefSiddon, vol_nc, s, t, l)
    renderer_case("trilinear_nc_b4", RefTrilinear, vol_nc, s, t, l, fwd_kw=dict(n_points=160))
    renderer_case("siddon_nc_b4_max", RefSiddon, vol_nc, s, t, l, ctor_kw=dict(reducefn="max"), grads=False)
    renderer_case("trilinear_nc_b4_max", RefTrilinear, vol_nc, s, t, l, ctor_kw=dict(reducefn="max"),
                  fwd_kw=dict(n_points=96), grads=False)
    # Q6: stop_gradients_through_grid_sample (renderers.py:63-65)
    renderer_case("siddon_nc_b4_stopgrad", RefSiddon, vol_nc, s, t, l,
                  ctor_kw=dict(stop_gradients_through_grid_sample=True))
    # Q8: voxel_shift = 0 (renderers.py:97-99,152)
    renderer_case("siddon_nc_b4_shift0", RefSiddon, vol_nc, s, t, l, ctor_kw=dict(voxel_shift=0.0))
    renderer_case("trilinear_nc_b4_shift0", RefTrilinear, vol_nc, s, t, l, ctor_kw=dict(voxel_shift=0.0),
                  fwd_kw=dict(n_points=120))
    # Q3: explicit alphamin/alphamax (renderers.py:214-223)
    renderer_case("trilinear_nc_b4_alpha", RefTrilinear, vol_nc, s, t, l,
                  fwd_kw=dict(n_points=100, alphamin=0.62, alphamax=0.97))
    # align_corners=True (Q8, renderers.py:40,161)
    renderer_case("siddon_nc_b4_ac", RefSiddon, vol_nc, s, t, l, fwd_kw=dict(align_corners=True), grads=False)
    renderer_case("trilinear_nc_b4_ac", RefTrilinear, vol_nc, s, t, l, fwd_kw=dict(n_points=90, align_corners=True),
                  grads=False)
    # Q10: ragged ray subset (odd N, not a grid)
    g = torch.Generator().manual_seed(7)
    pick = torch.randperm(t.shape[1], generator=g)[:57]
    renderer_case("siddon_nc_b4_ragged", RefSiddon, vol_nc, s, t[:, pick].contiguous(), l[:, :, pick].contiguous())
    renderer_case("trilinear_nc_b4_ragged", RefTrilinear, vol_nc, s, t[:, pick].contiguous(),
                  l[:, :, pick].contiguous(), fwd_kw=dict(n_points=77))

    # Q1: source (and for one pose the detector) inside the volume's box -> the infinite line is integrated
    rot_in = torch.tensor([[0.0, 0.0, 0.0], [0.3, -0.2, 0.1]])
    xyz_in = torch.tensor([[0.0, 90.0, 0.0], [10.0, 60.0, -20.0]])
    s1, t1, l1 = ref_rays(vol_nc, 14, 14, rot_in, xyz_in)
    renderer_case("siddon_nc_inside", RefSiddon, vol_nc, s1, t1, l1)
    renderer_case("trilinear_nc_inside", RefTrilinear, vol_nc, s1, t1, l1, fwd_kw=dict(n_points=150))

    # Q2: hand-made rays in voxel coordinates: axis-parallel, along a voxel face, missing the volume,
    # zero-length direction components, grazing a corner
    src = torch.tensor([[[5.3, 7.1, -20.0]], [[-15.0, 11.0, 13.0]]])
    tgt = torch.tensor([
        [[5.3, 7.1, 60.0], [5.3, 30.0, 60.0], [5.5, 7.5, 60.0], [100.0, 7.1, -20.0], [23.2, 31.1, 39.4],
         [-0.5, -0.5, 39.5], [5.3, 7.1, -19.0]],
        [[40.0, 11.0, 13.0], [40.0, 11.5, 13.5], [40.0, 50.0, 13.0], [-15.0, 11.0, 50.0], [23.5, 31.5, 39.5],
         [12.0, 16.0, 20.0], [-14.0, 11.2, 13.1]],
    ])
    ln = (tgt - src).norm(dim=-1).unsqueeze(1)
    renderer_case("siddon_nc_axis", RefSiddon, vol_nc, src, tgt, ln)
    renderer_case("trilinear_nc_axis", RefTrilinear, vol_nc, src, tgt, ln, fwd_kw=dict(n_points=200))

    # mask_to_channels (renderers.py:77-89, 242-252): per-label channels from a label volume
    i0, i1, i2 = np.meshgrid(np.arange(24), np.arange(32), np.arange(40), indexing="ij")
    labels = ((i0 * 3) // 24 + 2 * ((i2 * 2) // 40) + ((i1 > 20) & (i0 > 10))).astype(np.float32)  # values 0..4
    np.savez_compressed(os.path.join(HERE, "labels_nc.npz"), labels=labels)
    mask_t = torch.from_numpy(labels)
    for tag, cls, fkw in (("siddon_nc_b4_mask", RefSiddon, {}), ("trilinear_nc_b4_mask", RefTrilinear, dict(n_points=110))):
        rec = {"source": s.numpy(), "target": t.numpy(), "raylen": l.numpy(), "volume_key": np.str_("nc")}
        for dt_tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            out = cls()(torch.from_numpy(vol_nc).to(dt), s.to(dt), t.to(dt), l.to(dt), mask=mask_t.to(dt), **fkw)
            rec["img_" + dt_tag] = out.numpy()
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **rec)
        print(f"{tag}: shape {rec['img_f64'].shape} f32-vs-f64 "
              f"{np.abs(rec['img_f32'] - rec['img_f64']).max() / np.abs(rec['img_f64']).max():.2e}")

    # ---- BASELINE config[0]: 64^3 -> 64^2, 1 pose (volume regenerated from its seed at test time) -------
    vol64 = synthetic.make_volume(64, "rand", seed=0)
    rot1, xyz1 = synthetic.make_poses(1)
    s, t, l = ref_rays(vol64, 64, 64, rot1, xyz1)
    renderer_case("siddon_c1", RefSiddon, vol64, s, t, l, grads=False, store_vol="rand64")
    renderer_case("trilinear_c1", RefTrilinear, vol64, s, t, l, grads=False, store_vol="rand64")

    # ---- DRR-level (pose in, image out, pose gradients through convert -> Detector -> renderer) ---------
    vol32 = synthetic.make_volume(32, "phantom", seed=2)
    drr_case("drr_siddon_b4", vol32, 24, rot4, xyz4, "siddon")
    drr_case("drr_trilinear_b4", vol32, 24, rot4, xyz4, "trilinear", fwd_kw=dict(n_points=200))
    drr_case("drr_siddon_b1_patch", vol32, 24, rot1, xyz1, "siddon", patch_size=12)


if __name__ == "__main__":
    main()
