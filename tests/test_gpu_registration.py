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
