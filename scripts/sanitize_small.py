#!/usr/bin/env python
"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck): brick-major TMA forward
(several bricks per CTA, partial bricks, two pose chunks), the locality-ordered slab kernels, PeerGather is multi-GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffdrr_b200 import DRR, Siddon, _lib, renderers, synthetic  # noqa: E402
from diffdrr_b200.pose import convert  # noqa: E402
from diffdrr_b200.renderers import _ptr, _stream  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
for dims, H, B in (((50, 72, 64), 40, 3), ((30, 40, 36), 24, 35)):
    vol = torch.rand(*dims, device=dev)
    subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1))
    subj.volume.affine = synthetic.make_affine(dims)
    drr = DRR(subj, **synthetic.detector_kwargs(H)).to(dev)
    rot, xyz = synthetic.make_poses(B, seed=1)
    with torch.no_grad():
        src, tgt = drr.detector(convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY"), None)
        raylen = (tgt - src).norm(dim=-1).reshape(B, -1).contiguous()
        src, tgt = drr.affine_inverse(src).reshape(B, 3).contiguous(), drr.affine_inverse(tgt).contiguous()
    N = H * H
    ref = torch.empty(B, N, device=dev)
    _lib.check(lib.b200drr_siddon_fwd(_ptr(vol), *dims, _ptr(src), _ptr(tgt), _ptr(raylen), _ptr(ref), B, N, 0.5, 1e-8, 0, 0, _stream()), "fwd")
    ws = torch.empty(lib.b200drr_siddon_brick_workspace_bytes(B, H, H), dtype=torch.uint8, device=dev)
    for v in (0, 7):
        out = torch.zeros(B, N, device=dev)
        _lib.check(lib.b200drr_siddon_fwd_brick(_ptr(vol), *dims, _ptr(src), _ptr(tgt), _ptr(raylen), None, None, None, None, _ptr(out),
                                                _ptr(ws), ws.numel(), B, H, H, 0.5, 1e-8, v, _stream()), "brick")
        torch.cuda.synchronize()
        print(dims, "brick variant", v, "maxdiff", float((out - ref).abs().max() / ref.abs().max()))
    renderers._SORT_MIN_RAYS = 1
    s3, t3, l3 = src.reshape(B, 1, 3).clone().requires_grad_(True), tgt.clone().requires_grad_(True), raylen.reshape(B, 1, N)
    o = Siddon()(vol, s3, t3, l3)
    o.sum().backward()
    torch.cuda.synchronize()
    print(dims, "sorted sens maxdiff", float((o.detach().reshape(B, N) - ref).abs().max() / ref.abs().max()))
print("done")
