"""CPU checks of the path's callers: NCC loss and the Registration wrapper."""
import torch

from diffdrr_b200 import NormalizedCrossCorrelation2d, Registration, synthetic
from diffdrr_b200.drr import DRR


def _ncc_reference(x1, x2, eps=1e-5):
    """Formula of reference diffdrr/metrics.py:29-44 (patch_size=None), restated for the test."""
    def norm(x):
        mu = x.mean(dim=[-1, -2], keepdim=True)
        var = x.var(dim=[-1, -2], keepdim=True, correction=0) + eps
        return (x - mu) / var.sqrt()
    _, c, h, w = x1.shape
    return torch.einsum("b...,b...->b", norm(x1), norm(x2)) / (c * h * w)


def test_ncc_matches_reference_formula_and_bounds():
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.randn(5, 1, 17, 23, generator=g), torch.randn(5, 1, 17, 23, generator=g)
    ncc = NormalizedCrossCorrelation2d()
    assert torch.allclose(ncc(x1, x2), _ncc_reference(x1, x2), atol=1e-6)
    assert torch.allclose(ncc(x1, 3.0 * x1 + 2.0), torch.ones(5), atol=1e-3)      # invariant to affine intensity maps
    assert torch.allclose(ncc(x1, -x1), -torch.ones(5), atol=1e-3)
    x1.requires_grad_(True)
    ncc(x1, x2).sum().backward()
    assert torch.isfinite(x1.grad).all()


def test_registration_module_holds_pose_parameters():
    vol = synthetic.make_volume(8, "smooth")
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(6))
    rot, xyz = synthetic.make_poses(1)
    reg = Registration(drr, rot.clone(), xyz.clone(), "euler_angles", "ZXY")
    names = dict(reg.named_parameters())
    assert set(names) == {"_rotation", "_translation"} and reg.rotation is names["_rotation"]
    pose = reg.pose
    assert pose.matrix.shape == (1, 4, 4) and pose.matrix.requires_grad
    # camera centre = R @ t (reference pose.py:149-157): for rot = 0 it is the translation itself
    assert torch.allclose(pose.matrix[0, :3, 3], xyz[0])


def test_ncc_matches_the_reference_class_values_and_gradients():
    """NormalizedCrossCorrelation2d against goldens recorded from the UNMODIFIED reference class (metrics.py:21-44, kornia
    stubbed in tests/golden/make_golden_ncc.py): scores and d(score)/d(x2), full-image and patch mode (patch_size=5)."""
    import os

    import numpy as np

    from conftest import GOLDEN, relerr
    g = dict(np.load(os.path.join(GOLDEN, "ncc_reference.npz")))
    for tag, patch in (("full", None), ("patch5", 5)):
        for dt, suf, tol in ((torch.float32, "f32", 1e-5), (torch.float64, "f64", 1e-12)):
            x1 = torch.as_tensor(g["x1"]).to(dt)
            x2 = torch.as_tensor(g["x2"]).to(dt).requires_grad_(True)
            score = NormalizedCrossCorrelation2d(patch_size=patch)(x1, x2)
            (score * torch.tensor([1.0, -2.0, 0.5], dtype=dt)).sum().backward()
            assert relerr(score.detach().numpy(), g[f"{tag}_score_{suf}"]) < tol, (tag, suf)
            assert relerr(x2.grad.numpy(), g[f"{tag}_grad_x2_{suf}"]) < max(tol, 1e-5 if dt == torch.float32 else tol), (tag, suf)


def test_ncc_kernel_math_matches_the_reference_class_and_autograd():
    """ncc.cu's math (ncc_math.cuh compiled for the CPU by tests/hostemu: one-pass double moments + closed-form gradient) against
    (a) the goldens of the UNMODIFIED reference class and (b) torch autograd of the reference formula in fp64, for several
    channels, image sizes with a ragged last chunk, a large intensity offset (the one-pass variance must not cancel) and
    gradients with respect to BOTH images."""
    import os

    import numpy as np

    from conftest import GOLDEN, relerr
    from hostemu import emu
    g = dict(np.load(os.path.join(GOLDEN, "ncc_reference.npz")))
    w = np.array([1.0, -2.0, 0.5], np.float32)
    score, _, g2 = emu.ncc(g["x1"], g["x2"], w)
    assert relerr(score, g["full_score_f64"]) < 2e-6
    assert relerr(g2, g["full_grad_x2_f64"]) < max(1e-5, 2 * relerr(g["full_grad_x2_f32"], g["full_grad_x2_f64"]))
    gen = torch.Generator().manual_seed(3)
    for (B, C, H, W), offset in (((2, 3, 50, 47), 0.0), ((1, 1, 64, 64), 0.0), ((2, 1, 33, 3), 500.0), ((1, 2, 1, 5), 0.0)):
        x1 = torch.rand(B, C, H, W, generator=gen) * 3.0 + offset
        x2 = (0.5 * x1 + torch.rand(B, C, H, W, generator=gen)).contiguous()
        wb = torch.rand(B, generator=gen) - 0.3
        a = x1.double().requires_grad_(True)
        b = x2.double().requires_grad_(True)
        ref = _ncc_reference(a, b)
        (ref * wb.double()).sum().backward()
        score, g1, g2 = emu.ncc(x1.numpy(), x2.numpy(), wb.numpy())
        assert relerr(score, ref.detach().numpy()) < 5e-6, (B, C, H, W)
        assert relerr(g1, a.grad.numpy()) < 2e-4 and relerr(g2, b.grad.numpy()) < 2e-4, (B, C, H, W, relerr(g2, b.grad.numpy()))
    # a constant image: variance 0, the score is 0 and nothing is NaN (eps keeps the normalisation finite, as in the reference)
    score, g1, g2 = emu.ncc(np.full((1, 1, 8, 8), 2.0, np.float32), np.random.default_rng(0).random((1, 1, 8, 8), np.float32),
                            np.ones(1, np.float32))
    assert np.isfinite(score).all() and np.isfinite(g1).all() and np.isfinite(g2).all() and abs(float(score[0])) < 1e-3


def test_cached_detector_axes_follow_the_detector():
    """DRR._grid_axes: the pixel-row / pixel-column coordinates handed to the pose-in kernels are cached contiguous copies of
    detector.target[:, 0, 1] / [0, :, 0]; a row block (ray sharding) and a detector rebuilt by set_intrinsics_ get their own."""
    vol = synthetic.make_volume((20, 24, 28), "phantom", seed=3)
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(32))
    grid = drr.detector.target.view(32, 32, 3)
    rows, cols = drr._grid_axes(None)
    assert torch.equal(rows, grid[:, 0, 1]) and torch.equal(cols, grid[0, :, 0]) and rows.is_contiguous() and cols.is_contiguous()
    assert drr._grid_axes(None)[0] is rows                     # cached
    block, _ = drr._grid_axes((2, 5))
    assert torch.equal(block, grid[2:5, 0, 1])
    drr.set_intrinsics_(height=16, width=20)
    rows2, cols2 = drr._grid_axes(None)
    assert rows2.numel() == 16 and cols2.numel() == 20
