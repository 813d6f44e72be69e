"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes (no GPU, no kernels)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from diffdrr_b200.parallel import all_gather_images, global_alpha_range, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 16, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, batch):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(batch, rank, world)
        full = torch.arange(batch * 6, dtype=torch.float32).reshape(batch, 1, 2, 3)
        local = full[lo:hi].clone().requires_grad_(True)
        out = all_gather_images(local * 2.0, batch)
        assert torch.equal(out, full * 2.0), "gathered stack differs from the unsharded one"
        # backward: only this rank's slice of the upstream gradient comes back
        w = torch.linspace(0, 1, out.numel()).reshape(out.shape)
        (out * w).sum().backward()
        assert torch.allclose(local.grad, 2.0 * w[lo:hi])
        # trilinear sampling range: MIN / MAX over ranks
        a = torch.tensor(0.3 + 0.1 * rank, requires_grad=True)
        b = torch.tensor(0.8 + 0.05 * rank, requires_grad=True)
        amin, amax = global_alpha_range(a, b)
        assert abs(float(amin) - 0.3) < 1e-7 and abs(float(amax) - (0.8 + 0.05 * (world - 1))) < 1e-7
        # backward: every rank contributes (rank+1)*d/d amin and 10*(rank+1)*d/d amax; the owner gets the SUM
        ((rank + 1.0) * amin + 10.0 * (rank + 1.0) * amax).backward()
        tot = sum(r + 1.0 for r in range(world))
        assert float(a.grad) == (tot if rank == 0 else 0.0)
        assert float(b.grad) == (10.0 * tot if rank == world - 1 else 0.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_gather_and_range_world2(batch):
    port = 29500 + (os.getpid() % 2000) + batch
    mp.spawn(_worker, args=(2, port, batch), nprocs=2, join=True)


class _StubRenderer(torch.nn.Module):
    """CPU stand-in for the CUDA renderers: a smooth, per-ray-independent function of (source, target, raylen) plus the
    batch-global term Trilinear has (alphamin/alphamax kwargs), so that sharding mistakes change the result."""

    voxel_shift, eps = 0.5, 1e-8
    detector_shape = None

    def forward(self, volume, source, target, img, alphamin=None, alphamax=None, **kw):
        d = (target - source) * 1e-2
        val = torch.sin(d).sum(-1) + 0.1 * torch.cos(target * 1e-2).prod(-1)
        out = val.unsqueeze(1) * img * 1e-3
        if alphamin is not None:
            out = out * (1.0 + alphamin) + alphamax
        return out


def _ray_worker(rank, world, port, batch, height):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffdrr_b200 import DRR, synthetic
        from diffdrr_b200.parallel import render_sharded
        vol = torch.zeros(8, 8, 8)
        drr = DRR(synthetic.make_subject(vol), sdd=1020.0, height=height, width=6, delx=4.0)
        drr.renderer = _StubRenderer()
        rot0, xyz0 = synthetic.make_poses(batch, seed=3)
        w = torch.linspace(0.5, 1.5, batch * height * 6).reshape(batch, 1, height, 6)
        kw = dict(parameterization="euler_angles", convention="ZXY")
        rot, xyz = rot0.clone().requires_grad_(True), xyz0.clone().requires_grad_(True)
        ref = drr(rot, xyz, **kw)
        (ref * w).sum().backward()
        r2, x2 = rot0.clone().requires_grad_(True), xyz0.clone().requires_grad_(True)
        out = render_sharded(drr, r2, x2, shard="rays", **kw)
        assert out.shape == ref.shape and torch.allclose(out, ref, rtol=1e-6, atol=1e-7), "ray-sharded image differs"
        (out * w).sum().backward()
        # every rank ends up with the FULL pose gradient (partials summed in backward)
        assert torch.allclose(r2.grad, rot.grad, rtol=1e-4, atol=1e-6) and torch.allclose(x2.grad, xyz.grad, rtol=1e-4, atol=1e-7)
        # "auto" picks rays when there are fewer poses than ranks, poses otherwise
        auto = render_sharded(drr, rot0, xyz0, **kw)
        assert torch.allclose(auto, ref.detach(), rtol=1e-6, atol=1e-7)
        # local block only
        blk = render_sharded(drr, rot0, xyz0, shard="rays", gather=False, **kw)
        lo, hi = shard_bounds(height, rank, world)
        assert torch.allclose(blk, ref.detach()[:, :, lo:hi], rtol=1e-6, atol=1e-7)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch,height", [(1, 7), (3, 8)])
def test_ray_sharding_world2(batch, height):
    """Detector rows split over 2 gloo ranks (ragged when H is odd): gathered image == unsharded, pose gradients summed."""
    port = 31500 + (os.getpid() % 2000) + batch
    mp.spawn(_ray_worker, args=(2, port, batch, height), nprocs=2, join=True)


def test_balanced_pose_assignment():
    from diffdrr_b200.parallel import balanced_pose_assignment
    import random
    rnd = random.Random(0)
    costs = [rnd.uniform(0.8, 1.2) for _ in range(128)]
    groups = balanced_pose_assignment(costs, 8)
    assert sorted(i for g in groups for i in g) == list(range(128)) and all(len(g) == 16 for g in groups)
    loads = [sum(costs[i] for i in g) for g in groups]
    contiguous = [sum(costs[r * 16:(r + 1) * 16]) for r in range(8)]
    assert max(loads) / (sum(loads) / 8) < 1.005 < max(contiguous) / (sum(contiguous) / 8)
    import pytest
    with pytest.raises(ValueError):
        balanced_pose_assignment(costs[:10], 8)
