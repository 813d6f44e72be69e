"""Parity of the CUDA path (through the C ABI) with the unmodified reference's recorded outputs and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from gpu_common import DEV, grad_tol, t
from test_oracle import SIDDON, TRILINEAR

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4  # north_star: <= 1e-4 relative error vs the reference (max-abs / max-ref)


def _siddon(kw):
    from diffdrr_b200 import Siddon
    ctor = {k: v for k, v in kw.items() if k in ("voxel_shift", "stop_gradients_through_grid_sample")}
    if "reduce" in kw:
        ctor["reducefn"] = kw["reduce"]
    if kw.get("stop_grad"):
        ctor["stop_gradients_through_grid_sample"] = True
    return Siddon(**ctor), {k: v for k, v in kw.items() if k == "align_corners"}


def _trilinear(kw):
    from diffdrr_b200 import Trilinear
    ctor = {k: v for k, v in kw.items() if k == "voxel_shift"}
    if "reduce" in kw:
        ctor["reducefn"] = kw["reduce"]
    return Trilinear(**ctor), {k: v for k, v in kw.items() if k in ("n_points", "align_corners", "alphamin", "alphamax")}


@pytest.mark.parametrize("name,kw", SIDDON)
def test_siddon_forward_golden(name, kw):
    g = load_golden(name)
    mod, fkw = _siddon(kw)
    out = mod(t(g["volume"]), t(g["source"]), t(g["target"]), t(g["raylen"]), **fkw).cpu().numpy()
    assert out.shape == g["img_f32"].shape
    assert relerr(out, g["img_f32"]) < IMG_TOL
    assert relerr(out, g["img_f64"]) < IMG_TOL


@pytest.mark.parametrize("name,kw", TRILINEAR)
def test_trilinear_forward_golden(name, kw):
    g = load_golden(name)
    mod, fkw = _trilinear(kw)
    out = mod(t(g["volume"]), t(g["source"]), t(g["target"]), t(g["raylen"]), **fkw).cpu().numpy()
    assert relerr(out, g["img_f32"]) < IMG_TOL
    assert relerr(out, g["img_f64"]) < IMG_TOL


@pytest.mark.parametrize("name,kw", [
    ("siddon_nc_b4", {}), ("siddon_nc_b4_shift0", dict(voxel_shift=0.0)), ("siddon_nc_b4_ragged", {}),
    ("siddon_nc_b4_stopgrad", dict(stop_grad=True)),
])
def test_siddon_backward_golden(name, kw):
    g = load_golden(name)
    mod, fkw = _siddon(kw)
    v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    (mod(v, s, tg, l, **fkw) * t(g["w"])).sum().backward()
    assert relerr(tg.grad.cpu().numpy(), g["g_target_f64"]) < grad_tol(g, "g_target")
    assert relerr(s.grad.cpu().numpy(), g["g_source_f64"]) < grad_tol(g, "g_source")
    if kw.get("stop_grad"):
        assert v.grad is None or not v.grad.any()
        assert l.grad is None or not l.grad.any()
    else:
        assert relerr(l.grad.cpu().numpy(), g["g_raylen_f64"]) < grad_tol(g, "g_raylen")
        assert relerr(v.grad.cpu().numpy(), g["g_volume_f64"]) < grad_tol(g, "g_volume")


@pytest.mark.parametrize("name,kw", [
    ("siddon_nc_b4", {}), ("siddon_nc_b4_shift0", dict(voxel_shift=0.0)), ("siddon_nc_b4_ragged", {}),
    ("siddon_nc_b4_stopgrad", dict(stop_grad=True)), ("siddon_nc_inside", {}),
])
def test_siddon_ray_gradients_only_golden(name, kw):
    """Only the ray tensors require grad (registration): the binding takes the one-walk forward-with-sensitivities entry
    (b200drr_siddon_fwd_sens for arbitrary ray sets) + the elementwise backward.  Same goldens, same bar."""
    g = load_golden(name)
    mod, fkw = _siddon(kw)
    v, s, tg, l = t(g["volume"]), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    out = mod(v, s, tg, l, **fkw)
    assert relerr(out.detach().cpu().numpy(), g["img_f64"]) < 1e-4
    (out * t(g["w"])).sum().backward()
    assert relerr(tg.grad.cpu().numpy(), g["g_target_f64"]) < grad_tol(g, "g_target")
    assert relerr(s.grad.cpu().numpy(), g["g_source_f64"]) < grad_tol(g, "g_source")
    if kw.get("stop_grad"):
        assert l.grad is None or not l.grad.any()
    else:
        assert relerr(l.grad.cpu().numpy(), g["g_raylen_f64"]) < grad_tol(g, "g_raylen")


@pytest.mark.parametrize("name,kw", [
    ("trilinear_nc_b4", dict(n_points=160)),
    ("trilinear_nc_b4_alpha", dict(n_points=100, alphamin=0.62, alphamax=0.97)),
    ("trilinear_nc_b4_ragged", dict(n_points=77)),
    ("trilinear_nc_b4_shift0", dict(n_points=120, voxel_shift=0.0)),
])
def test_trilinear_backward_golden(name, kw):
    """Full autograd chain incl. the arg-min/arg-max branch of the batch-global alpha range (quirk Q3)."""
    g = load_golden(name)
    mod, fkw = _trilinear(kw)
    v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    (mod(v, s, tg, l, **fkw) * t(g["w"])).sum().backward()
    # trilinear pose gradients are fp32 sums of voxel differences: the reference's own fp32 run is 8e-4..1e-2 off
    # its fp64 run, so the floor is 1e-3
    assert relerr(tg.grad.cpu().numpy(), g["g_target_f64"]) < grad_tol(g, "g_target", 1e-3)
    assert relerr(s.grad.cpu().numpy(), g["g_source_f64"]) < grad_tol(g, "g_source", 1e-3)
    assert relerr(l.grad.cpu().numpy(), g["g_raylen_f64"]) < grad_tol(g, "g_raylen")
    assert relerr(v.grad.cpu().numpy(), g["g_volume_f64"]) < grad_tol(g, "g_volume")


@pytest.mark.parametrize("name,kw", [
    ("trilinear_nc_b4", dict(n_points=160)),
    ("trilinear_nc_b4_alpha", dict(n_points=100, alphamin=0.62, alphamax=0.97)),
    ("trilinear_nc_b4_ragged", dict(n_points=77)),
    ("trilinear_nc_b4_shift0", dict(n_points=120, voxel_shift=0.0)),
])
def test_trilinear_ray_gradients_only_golden(name, kw):
    """Only the ray tensors require grad: one march (b200drr_trilinear_fwd_sens, arbitrary ray set) + elementwise
    backward, incl. the arg-min/arg-max branch of the batch-global alpha range.  Same goldens, same bar."""
    g = load_golden(name)
    mod, fkw = _trilinear(kw)
    v, s, tg, l = t(g["volume"]), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    out = mod(v, s, tg, l, **fkw)
    assert relerr(out.detach().cpu().numpy(), g["img_f64"]) < 1e-4
    (out * t(g["w"])).sum().backward()
    assert relerr(tg.grad.cpu().numpy(), g["g_target_f64"]) < grad_tol(g, "g_target", 1e-3)
    assert relerr(s.grad.cpu().numpy(), g["g_source_f64"]) < grad_tol(g, "g_source", 1e-3)
    assert relerr(l.grad.cpu().numpy(), g["g_raylen_f64"]) < grad_tol(g, "g_raylen")


@pytest.mark.parametrize("name,renderer,fkw", [
    ("drr_siddon_b4", "siddon", {}), ("drr_trilinear_b4", "trilinear", dict(n_points=200)),
    ("drr_siddon_b1_patch", "siddon", {}),
])
def test_drr_module_golden(name, renderer, fkw):
    """Pose in, image out, pose gradients through convert -> Detector -> affine_inverse -> CUDA renderer."""
    from diffdrr_b200 import DRR, synthetic
    g = load_golden(name)
    H = int(g["height"])
    extra = dict(patch_size=12) if name.endswith("patch") else {}
    drr = DRR(synthetic.make_subject(g["volume"]), **synthetic.detector_kwargs(H), renderer=renderer, **extra).to(DEV)
    rot, xyz = t(g["rot"], True), t(g["xyz"], True)
    img = drr(rot, xyz, parameterization="euler_angles", convention="ZXY", **fkw)
    assert tuple(img.shape) == g["img_f32"].shape
    assert relerr(img.detach().cpu().numpy(), g["img_f64"]) < IMG_TOL
    if name.endswith("patch"):
        return  # canonical axis-aligned pose: the reference's own fp32 pose gradient is off by O(1) there
    (img * t(g["w"])).sum().backward()
    floor = 1e-3 if renderer == "trilinear" else 1e-4
    assert relerr(rot.grad.cpu().numpy(), g["g_rot_f64"]) < grad_tol(g, "g_rot", floor)
    assert relerr(xyz.grad.cpu().numpy(), g["g_xyz_f64"]) < grad_tol(g, "g_xyz", floor)


def test_mask_to_channels_golden():
    """Per-label channels (reference renderers.py:77-89, 242-252) through the modules and through DRR(mask_to_channels=True)."""
    import os
    from conftest import GOLDEN
    from diffdrr_b200 import Siddon, Trilinear
    labels = np.load(os.path.join(GOLDEN, "labels_nc.npz"))["labels"]
    for name, mod, fkw in (("siddon_nc_b4_mask", Siddon(), {}), ("trilinear_nc_b4_mask", Trilinear(), dict(n_points=110))):
        g = load_golden(name)
        with torch.no_grad():
            out = mod(t(g["volume"]), t(g["source"]), t(g["target"]), t(g["raylen"]), mask=t(labels), **fkw)
        assert tuple(out.shape) == g["img_f64"].shape
        assert relerr(out.cpu().numpy(), g["img_f64"]) < IMG_TOL
    # through the DRR module: channels sum to the plain DRR
    from diffdrr_b200 import DRR, synthetic
    g = load_golden("siddon_nc_b4")
    subj = synthetic.make_subject(g["volume"], mask=torch.from_numpy(labels))
    drr = DRR(subj, **synthetic.detector_kwargs(18)).to(DEV)
    rot, xyz = synthetic.make_poses(4, seed=0)
    with torch.no_grad():
        ch = drr(rot.to(DEV), xyz.to(DEV), parameterization="euler_angles", convention="ZXY", mask_to_channels=True)
        full = drr(rot.to(DEV), xyz.to(DEV), parameterization="euler_angles", convention="ZXY")
    assert ch.shape == (4, 6, 18, 18)
    assert relerr(ch.sum(1, keepdim=True).cpu().numpy(), full.cpu().numpy()) < 2e-5


def test_mask_to_channels_backward_golden():
    """Autograd through mask_to_channels rendering (b200drr_*_bwd_mask) against the reference's own autograd of the
    scatter_add_ routing (tests/golden/make_golden_mask_grads.py): loss = sum(w * img), img (B, C, N)."""
    import os
    from conftest import GOLDEN
    from diffdrr_b200 import Siddon, Trilinear
    labels = np.load(os.path.join(GOLDEN, "labels_nc.npz"))["labels"]
    for name, mod, fkw, floor in (("siddon_nc_b4_mask", Siddon(), {}, 1e-4),
                                  ("trilinear_nc_b4_mask", Trilinear(), dict(n_points=110), 1e-3)):
        g = load_golden(name)
        gg = np.load(os.path.join(GOLDEN, name + "_grad.npz"))
        v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
        out = mod(v, s, tg, l, mask=t(labels), **fkw)
        assert relerr(out.detach().cpu().numpy(), g["img_f64"]) < IMG_TOL
        (out * t(gg["w"])).sum().backward()
        for key, x in (("g_target", tg), ("g_source", s), ("g_raylen", l), ("g_volume", v)):
            tol = max(floor if key in ("g_target", "g_source") else 1e-4, 2.0 * relerr(gg[key + "_f32"], gg[key + "_f64"]))
            assert relerr(x.grad.cpu().numpy(), gg[key + "_f64"]) < tol, (name, key)
    # stop_gradients_through_grid_sample: no volume / ray-length gradient, ray gradients unchanged
    g = load_golden("siddon_nc_b4_mask")
    gg = np.load(os.path.join(GOLDEN, "siddon_nc_b4_mask_grad.npz"))
    v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    (Siddon(stop_gradients_through_grid_sample=True)(v, s, tg, l, mask=t(labels)) * t(gg["w"])).sum().backward()
    assert v.grad is None and l.grad is None
    assert relerr(tg.grad.cpu().numpy(), gg["g_target_f64"]) < 1e-4


def test_option_backward_golden():
    """Autograd through the options outside the fast kernels -- Siddon align_corners=True, reducefn="max" on both
    renderers -- against the reference's own autograd (tests/golden/make_golden_extra_grads.py)."""
    import os
    from conftest import GOLDEN
    from diffdrr_b200 import Siddon, Trilinear
    cases = (("siddon_nc_b4_ac", Siddon(), dict(align_corners=True), 1e-4),
             ("siddon_nc_b4_max", Siddon(reducefn="max"), {}, 1e-4),
             ("trilinear_nc_b4_max", Trilinear(reducefn="max"), dict(n_points=96), 1e-3))
    for name, mod, fkw, floor in cases:
        g = load_golden(name)
        gg = np.load(os.path.join(GOLDEN, name + "_grad.npz"))
        v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
        out = mod(v, s, tg, l, **fkw)
        assert relerr(out.detach().cpu().numpy(), g["img_f64"]) < IMG_TOL
        (out * t(g["w"])).sum().backward()
        for key, x in (("g_target", tg), ("g_source", s), ("g_raylen", l), ("g_volume", v)):
            tol = max(floor if key in ("g_target", "g_source") else 1e-4, 2.0 * relerr(gg[key + "_f32"], gg[key + "_f64"]))
            assert relerr(x.grad.cpu().numpy(), gg[key + "_f64"]) < tol, (name, key)


@pytest.mark.parametrize("name,stop", [("siddon_nc_b4_bilinear", False), ("siddon_nc_b4_bilinear_stopgrad", True)])
def test_siddon_bilinear_mode_golden(name, stop):
    """Siddon(mode="bilinear") -- trilinear sampling at the segment midpoints -- image and autograd against the reference,
    with and without stop_gradients_through_grid_sample (the flag only matters in this mode: quirk Q6)."""
    from diffdrr_b200 import Siddon
    g = load_golden(name)
    mod = Siddon(mode="bilinear", stop_gradients_through_grid_sample=stop)
    v, s, tg, l = t(g["volume"], True), t(g["source"], True), t(g["target"], True), t(g["raylen"], True)
    out = mod(v, s, tg, l)
    assert relerr(out.detach().cpu().numpy(), g["img_f64"]) < IMG_TOL
    (out * t(g["w"])).sum().backward()
    assert relerr(tg.grad.cpu().numpy(), g["g_target_f64"]) < grad_tol(g, "g_target", 1e-3)
    assert relerr(s.grad.cpu().numpy(), g["g_source_f64"]) < grad_tol(g, "g_source", 1e-3)
    if stop:
        assert v.grad is None and l.grad is None
    else:
        assert relerr(l.grad.cpu().numpy(), g["g_raylen_f64"]) < grad_tol(g, "g_raylen")
        assert relerr(v.grad.cpu().numpy(), g["g_volume_f64"]) < grad_tol(g, "g_volume")
    with torch.no_grad():  # reducefn="max" forward: bounded by the sum on a non-negative volume
        mx = Siddon(mode="bilinear", reducefn="max")(v, s, tg, l)
    assert bool((mx <= out.detach() * (1 + 1e-5) + 1e-6).all())


def test_unsupported_options_raise():
    from diffdrr_b200 import Siddon, Trilinear
    g = load_golden("siddon_nc_axis")
    args = (t(g["volume"]), t(g["source"]), t(g["target"]), t(g["raylen"]))
    with pytest.raises(NotImplementedError):
        Siddon(filter_intersections_outside_volume=True)(*args)   # quirk Q5: the reference crashes too
    with pytest.raises(ValueError):
        Siddon(mode="bicubic")(*args)
    with pytest.raises(NotImplementedError):  # a callable reducefn (tests/test_gpu_callable.py) is not combined with a mask
        Siddon(reducefn=lambda x: x.mean(-1))(*args, mask=torch.zeros_like(args[0]))
    with pytest.raises(NotImplementedError):  # mask rendering: reducefn="sum", align_corners=False only
        Siddon()(*args, align_corners=True, mask=torch.zeros_like(args[0]))
    with pytest.raises(NotImplementedError):
        Siddon()(args[0].double(), *args[1:])


def test_mask_to_channels_tile_ordered_grid_kernels_equal_the_row_ordered_ones():
    """b200drr_*_fwd_mask_grid (pixel tiles; Siddon: slab-major too) vs b200drr_*_fwd_mask, on a grid that is no multiple of the
    tile and a volume of more than one slab."""
    from diffdrr_b200 import DRR, Siddon, Trilinear, synthetic
    vol = synthetic.make_volume((36, 44, 52), "phantom", seed=1)
    drr = DRR(synthetic.make_subject(vol), sdd=1020.0, height=21, width=30, delx=6.0).to(DEV)
    lab = (torch.arange(36, device=DEV)[:, None, None] // 13 + 3 * (torch.arange(52, device=DEV)[None, None, :] // 27)).float()
    lab = lab.expand(36, 44, 52).contiguous()
    rot, xyz = synthetic.make_poses(3, seed=4)
    from diffdrr_b200.pose import convert
    src, tgt = drr.detector(convert(rot.to(DEV), xyz.to(DEV), parameterization="euler_angles", convention="ZXY"), None)
    img = (tgt - src).norm(dim=-1).unsqueeze(1)
    s, tg = drr.affine_inverse(src), drr.affine_inverse(tgt)
    wimg = torch.rand(3, 6, 21 * 30, device=DEV, generator=torch.Generator(DEV).manual_seed(5))
    for mod, kw in ((Siddon(), {}), (Trilinear(), dict(n_points=90))):
        res = []
        for grid in (None, (21, 30)):
            mod.detector_shape = grid
            v, ss, tt = drr.density.detach().clone().requires_grad_(True), s.detach().clone().requires_grad_(True), \
                tg.detach().clone().requires_grad_(True)
            out = mod(v, ss, tt, img, mask=lab, **kw)
            (out * wimg).sum().backward()     # b200drr_*_bwd_mask vs b200drr_*_bwd_mask_grid
            res.append((out.detach(), v.grad, ss.grad, tt.grad))
        rows, tiles = res
        assert rows[0].shape == tiles[0].shape == (3, 6, 21 * 30)
        # trilinear: same per-ray code, threads re-ordered -> bitwise; Siddon: slab-major partial sums -> fp32 round-off
        assert relerr(tiles[0].cpu().numpy(), rows[0].cpu().numpy()) < (1e-5 if isinstance(mod, Siddon) else 1e-12)
        for a, b in zip(tiles[1:], rows[1:]):   # gradients: identical per-ray math, atomics / block sums in another order
            assert relerr(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
