"""Pose-in entry of the trilinear training path (b200drr_trilinear_alpha_range_pose / _fwd_sens_pose / _bwd_sens_pose; reference
detector.py:144-154 + drr.py:201-205 + renderers.py:205-240 collapsed): against the ray-tensor path it replaces."""
import pytest
import torch

from conftest import relerr
from gpu_common import DEV

pytestmark = pytest.mark.gpu


def _drr(shape=(40, 48, 56), det=36):
    from diffdrr_b200 import DRR, synthetic
    vol = synthetic.make_volume(shape, "phantom", seed=5)
    return DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(det), renderer="trilinear").to(DEV)


def _step(drr, rot, xyz, w, n_points, pose_in):
    import diffdrr_b200.drr as drr_mod
    keep = drr_mod._TRILINEAR_POSE_IN
    drr_mod._TRILINEAR_POSE_IN = pose_in
    try:
        r, x = rot.clone().requires_grad_(True), xyz.clone().requires_grad_(True)
        assert drr._pose_in_trilinear_ok(False, dict(n_points=n_points), (r, x)) == pose_in
        img = drr(r, x, parameterization="euler_angles", convention="ZXY", n_points=n_points)
        (img * w).sum().backward()
        return img.detach(), r.grad, x.grad
    finally:
        drr_mod._TRILINEAR_POSE_IN = keep


@pytest.mark.parametrize("B,n_points", [(3, 120), (1, 77), (12, 64)])
def test_pose_in_matches_the_ray_tensor_path(B, n_points):
    from diffdrr_b200 import synthetic
    drr = _drr()
    rot, xyz = synthetic.make_poses(max(B, 2), seed=2)
    rot, xyz = rot[:B].to(DEV), xyz[:B].to(DEV)
    w = torch.rand(B, 1, 36, 36, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    a = _step(drr, rot, xyz, w, n_points, True)
    b = _step(drr, rot, xyz, w, n_points, False)
    assert relerr(a[0].cpu().numpy(), b[0].cpu().numpy()) < 2e-5
    assert relerr(a[1].cpu().numpy(), b[1].cpu().numpy()) < 2e-3   # fp32 sums in different orders, ill-conditioned
    assert relerr(a[2].cpu().numpy(), b[2].cpu().numpy()) < 2e-3


def test_alpha_range_kernel_matches_get_alpha_minmax():
    """Values bit-close to the torch reduction over the materialised rays, and the returned rays attain them."""
    from diffdrr_b200 import _lib, geometry, synthetic
    from diffdrr_b200.pose import convert
    from diffdrr_b200.renderers import _get_alpha_minmax, _ptr, _stream
    drr = _drr((30, 44, 52), det=28)
    rot, xyz = synthetic.make_poses(5, seed=7)
    pose = convert(rot.to(DEV), xyz.to(DEV), parameterization="euler_angles", convention="ZXY")
    src, tgt = drr.detector(pose, None)
    s, t = drr.affine_inverse(src), drr.affine_inverse(tgt)
    dims = torch.tensor(drr.density.shape, device=DEV, dtype=torch.float32)
    amin, amax = _get_alpha_minmax(s, t, dims, 0.5, 1e-8)
    Q, r, Ainv = drr._pose_constants()
    src_v, G, Wd = geometry.pose_rays(pose.matrix, Q, r, Ainv)
    grid = drr.detector.target.view(28, 28, 3)
    rows, cols = grid[:, 0, 1].contiguous(), grid[0, :, 0].contiguous()
    rng, arg = torch.empty(2, device=DEV), torch.empty(2, dtype=torch.int64, device=DEV)
    scratch = torch.empty(2, dtype=torch.int64, device=DEV)
    _lib.check(_lib.load().b200drr_trilinear_alpha_range_pose(*drr.density.shape, _ptr(src_v.contiguous()), _ptr(G.contiguous()),
                                                              _ptr(Wd.contiguous()), _ptr(rows), _ptr(cols), _ptr(rng), _ptr(arg),
                                                              _ptr(scratch), 5, 28, 28, 0.5, 1e-8, _stream()), "alpha_range_pose")
    torch.cuda.synchronize()
    assert abs(float(rng[0]) - float(amin.min())) < 2e-6 and abs(float(rng[1]) - float(amax.max())) < 2e-6
    assert abs(float(amin.reshape(-1)[arg[0]]) - float(amin.min())) < 2e-6
    assert abs(float(amax.reshape(-1)[arg[1]]) - float(amax.max())) < 2e-6


def test_registration_of_a_trilinear_drr_takes_the_pose_in_path_and_converges():
    from diffdrr_b200 import synthetic
    from diffdrr_b200.metrics import NormalizedCrossCorrelation2d
    from diffdrr_b200.registration import Registration
    drr = _drr((48, 48, 48), det=32)
    true_rot, true_xyz = torch.zeros(1, 3, device=DEV), torch.tensor([[0.0, 850.0, 0.0]], device=DEV)
    with torch.no_grad():
        target = drr(true_rot, true_xyz, parameterization="euler_angles", convention="ZXY", n_points=96)
    reg = Registration(drr, true_rot + 0.05, true_xyz + 4.0, "euler_angles", "ZXY").to(DEV)
    assert drr._pose_in_trilinear_ok(False, dict(n_points=96), (reg.rotation, reg.translation))
    ncc = NormalizedCrossCorrelation2d()
    opt = torch.optim.Adam([{"params": [reg.rotation], "lr": 5e-3}, {"params": [reg.translation], "lr": 3e-1}])
    first = None
    for _ in range(120):
        opt.zero_grad()
        loss = 1.0 - ncc(target, reg(n_points=96)).mean()
        loss.backward()
        opt.step()
        first = float(loss) if first is None else first
    assert float(loss) < 0.25 * first
