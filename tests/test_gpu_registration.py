"""GPU tests of the path's caller (SURVEY.md 8f-1 / VERDICT r1 item 7): the 2D/3D registration loop of
notebooks/tutorials/registration.ipynb cell 10 -- Registration (registration.py:14-50) + NCC (metrics.py:21-44) + Adam --
through the CUDA kernels, eagerly and as one CUDA graph per step."""
import pytest
import torch

from gpu_common import DEV

pytestmark = pytest.mark.gpu


def _problem(D=128, H=96):
    from diffdrr_b200 import DRR, NormalizedCrossCorrelation2d, Registration, synthetic
    vol = synthetic.make_volume(D, "smooth", seed=1)
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(H), renderer="siddon",
              stop_gradients_through_grid_sample=True).to(DEV)
    true_rot, true_xyz = torch.tensor([[0.0, 0.0, 0.0]], device=DEV), torch.tensor([[0.0, 850.0, 0.0]], device=DEV)
    with torch.no_grad():
        target = drr(true_rot, true_xyz, parameterization="euler_angles", convention="ZXY")
    reg = Registration(drr, (true_rot + torch.tensor([[0.12, -0.08, 0.06]], device=DEV)).clone(),
                       (true_xyz + torch.tensor([[10.0, -20.0, 8.0]], device=DEV)).clone(), "euler_angles", "ZXY").to(DEV)
    return drr, reg, NormalizedCrossCorrelation2d(), target, true_rot, true_xyz


@pytest.mark.parametrize("graph", [False, True])
def test_registration_loop_converges(graph):
    drr, reg, ncc, target, true_rot, true_xyz = _problem()
    opt = torch.optim.Adam([{"params": [reg.rotation], "lr": 5e-3}, {"params": [reg.translation], "lr": 5e-1}], capturable=graph)

    def step():
        opt.zero_grad(set_to_none=False)
        loss = 1.0 - ncc(target, reg()).mean()
        loss.backward()
        opt.step()
        return loss

    with torch.no_grad():
        start = float(1.0 - ncc(target, reg()).mean())
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        run = g.replay
    else:
        run = step
    for _ in range(400):
        run()
    torch.cuda.synchronize()
    with torch.no_grad():
        final = float(1.0 - ncc(target, reg()).mean())
    assert start > 0.02 and final < 1e-3 and final < 0.05 * start
    assert float((reg.rotation.detach() - true_rot).abs().max()) < 0.02
    assert float((reg.translation.detach() - true_xyz).abs().max()) < 2.0


def test_ncc_gradient_flows_to_the_pose_through_the_kernels():
    drr, reg, ncc, target, _, _ = _problem(D=64, H=48)
    loss = 1.0 - ncc(target, reg()).mean()
    loss.backward()
    for p in (reg.rotation, reg.translation):
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0
    # patch-mode NCC (metrics.py:16-19) runs on the GPU images too
    from diffdrr_b200 import NormalizedCrossCorrelation2d
    s = NormalizedCrossCorrelation2d(patch_size=9)(target, reg().detach())
    assert s.shape == (1,) and -1.0 <= float(s) <= 1.0


def _ncc_torch(x1, x2, eps=1e-5):
    """The reference formula (metrics.py:29-44, patch_size None) in plain torch ops, for comparison with the CUDA kernels."""
    def norm(x):
        mu = x.mean(dim=(-1, -2), keepdim=True)
        return (x - mu) / (x.var(dim=(-1, -2), keepdim=True, correction=0) + eps).sqrt()
    return (norm(x1) * norm(x2)).flatten(1).sum(dim=1) / x1[0].numel()


def test_ncc_kernels_match_the_reference_goldens():
    """b200drr_ncc_fwd / _bwd (csrc/ncc.cu) against scores + d(score)/d(x2) recorded from the UNMODIFIED reference class
    (tests/golden/make_golden_ncc.py; N = 480 pixels: a single ragged chunk, vector path)."""
    import os

    import numpy as np

    from conftest import GOLDEN, relerr
    from diffdrr_b200 import NormalizedCrossCorrelation2d
    g = dict(np.load(os.path.join(GOLDEN, "ncc_reference.npz")))
    x1 = torch.as_tensor(g["x1"]).to(DEV)
    x2 = torch.as_tensor(g["x2"]).to(DEV).requires_grad_(True)
    ncc = NormalizedCrossCorrelation2d()
    assert ncc.fused_ok(x1, x2)
    score = ncc(x1, x2)
    (score * torch.tensor([1.0, -2.0, 0.5], device=DEV)).sum().backward()
    assert relerr(score.detach().cpu().numpy(), g["full_score_f64"]) < 2e-6
    assert relerr(x2.grad.cpu().numpy(), g["full_grad_x2_f64"]) < max(1e-5, 2 * relerr(g["full_grad_x2_f32"], g["full_grad_x2_f64"]))


@pytest.mark.parametrize("shape,offset", [((2, 3, 50, 47), 0.0), ((1, 1, 256, 256), 0.0), ((3, 1, 33, 3), 500.0),
                                          ((1, 2, 1, 5), 0.0), ((2, 1, 300, 300), 40.0)])
def test_ncc_kernels_match_autograd_of_the_reference_formula(shape, offset):
    """Scores and the gradients with respect to BOTH images against fp64 autograd of the reference formula: several channels,
    sizes that are not a multiple of four (scalar path) or of the 2048-pixel chunk, a large intensity offset."""
    from conftest import relerr
    from diffdrr_b200 import NormalizedCrossCorrelation2d
    gen = torch.Generator().manual_seed(5)
    x1 = (torch.rand(*shape, generator=gen) * 3.0 + offset).to(DEV)
    x2 = (0.5 * x1.cpu() + torch.rand(*shape, generator=gen)).to(DEV)
    wb = (torch.rand(shape[0], generator=gen) - 0.3).to(DEV)
    a, b = x1.double().requires_grad_(True), x2.double().requires_grad_(True)
    ref = _ncc_torch(a, b)
    (ref * wb.double()).sum().backward()
    p, q = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    score = NormalizedCrossCorrelation2d()(p, q)
    (score * wb).sum().backward()
    assert relerr(score.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 5e-6
    assert relerr(p.grad.cpu().numpy(), a.grad.cpu().numpy()) < 2e-4
    assert relerr(q.grad.cpu().numpy(), b.grad.cpu().numpy()) < 2e-4
    # only one input needs a gradient (the registration loop: the target is fixed); unaligned views take the scalar path
    q2 = x2.clone().requires_grad_(True)
    NormalizedCrossCorrelation2d()(x1, q2).sum().backward()
    s1 = NormalizedCrossCorrelation2d()(x1, x2)
    assert torch.equal(s1, NormalizedCrossCorrelation2d()(x1, x2))  # fixed summation order: run-to-run identical
    assert q2.grad is not None and torch.isfinite(q2.grad).all()
