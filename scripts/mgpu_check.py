#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/mgpu_check.py : sharded rendering == single-GPU rendering (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from diffdrr_b200 import DRR, synthetic
from diffdrr_b200.parallel import render_sharded, shard_bounds

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
vol = synthetic.make_volume(96, "smooth", seed=7)  # smooth: pose gradients on hard edges are fp32-ill-conditioned
B = 6
rot, xyz = synthetic.make_poses(B, seed=2)
ok = True
for renderer, kw in (("siddon", {}), ("trilinear", dict(n_points=150))):
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(64), renderer=renderer).to(dev)
    w = torch.rand(B, 1, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    r1, x1 = rot.to(dev).requires_grad_(True), xyz.to(dev).requires_grad_(True)
    ref = drr(r1, x1, parameterization="euler_angles", convention="ZXY", **kw)
    (ref * w).sum().backward()
    r2, x2 = rot.to(dev).requires_grad_(True), xyz.to(dev).requires_grad_(True)
    out = render_sharded(drr, r2, x2, parameterization="euler_angles", convention="ZXY", **kw)
    (out * w).sum().backward()
    lo, hi = shard_bounds(B, rank, world)
    e_img = float((out - ref).abs().max() / ref.abs().max())
    e_gr = float((r2.grad[lo:hi] - r1.grad[lo:hi]).abs().max() / r1.grad.abs().max())
    e_gx = float((x2.grad[lo:hi] - x1.grad[lo:hi]).abs().max() / x1.grad.abs().max())
    outside = float(r2.grad[:lo].abs().sum() + r2.grad[hi:].abs().sum())
    good = e_img < 1e-5 and e_gr < 1e-3 and e_gx < 1e-3 and outside == 0.0
    ok &= good
    print(f"[rank {rank}] {renderer}: image err {e_img:.2e}, grad err rot {e_gr:.2e} xyz {e_gx:.2e}, grads outside shard {outside} -> {'OK' if good else 'FAIL'}", flush=True)
    # ray sharding (detector rows split, every rank ends with the FULL pose gradient): the B < #GPUs partitioning
    for nb in (1, 3):
        r3, x3 = rot[:nb].to(dev).requires_grad_(True), xyz[:nb].to(dev).requires_grad_(True)
        ref3 = drr(r3, x3, parameterization="euler_angles", convention="ZXY", **kw)
        (ref3 * w[:nb]).sum().backward()
        r4, x4 = rot[:nb].to(dev).requires_grad_(True), xyz[:nb].to(dev).requires_grad_(True)
        out3 = render_sharded(drr, r4, x4, shard="rays", parameterization="euler_angles", convention="ZXY", **kw)
        (out3 * w[:nb]).sum().backward()
        e_img = float((out3 - ref3).abs().max() / ref3.abs().max())
        e_gr = float((r4.grad - r3.grad).abs().max() / r3.grad.abs().max())
        e_gx = float((x4.grad - x3.grad).abs().max() / x3.grad.abs().max())
        good = e_img < 1e-5 and e_gr < 1e-3 and e_gx < 1e-3
        ok &= good
        print(f"[rank {rank}] {renderer} ray-sharded B={nb}: image err {e_img:.2e}, grad err rot {e_gr:.2e} xyz {e_gx:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
