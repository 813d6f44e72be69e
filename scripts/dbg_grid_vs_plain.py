import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from diffdrr_b200 import DRR, Siddon, synthetic
from diffdrr_b200.pose import convert
dev = torch.device("cuda:0")
D, H, B = 512, 256, 4
x = torch.linspace(-1, 1, D, device=dev)
smooth = torch.exp(-(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2) / 0.3)
subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1)); subj.volume.affine = synthetic.make_affine(D)
drr = DRR(subj, **synthetic.detector_kwargs(H)).to(dev)
rot, xyz = synthetic.make_poses(B, seed=0)
with torch.no_grad():
    src, tgt = drr.detector(convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY"), None)
    raylen = (tgt - src).norm(dim=-1).unsqueeze(1).contiguous()
    src, tgt = drr.affine_inverse(src).contiguous(), drr.affine_inverse(tgt).contiguous()
w = torch.rand(B, 1, H * H, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
res = []
for hint in (None, (H, H)):
    m = Siddon(); m.detector_shape = hint
    s, t, l = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True), raylen.clone().requires_grad_(True)
    (m(smooth, s, t, l) * w).sum().backward()
    res.append((s.grad.clone(), t.grad.clone()))
ds = (res[0][0] - res[1][0]).abs(); dt = (res[0][1] - res[1][1]).abs()
print("g_src max rel", float(ds.max() / res[0][0].abs().max()), "g_tgt max rel", float(dt.max() / res[0][1].abs().max()), "max|g_tgt|", float(res[0][1].abs().max()))
flat = dt.amax(-1).flatten(); top = torch.topk(flat, 8).indices
for i in top.tolist():
    b, n = divmod(i, H * H)
    print(b, n, "src", src[b, 0].tolist(), "tgt", tgt[b, n].tolist(), "w", float(w[b, 0, n]), "L", float(raylen[b, 0, n]), "plain", res[0][1][b, n].tolist(), "tiled", res[1][1][b, n].tolist())
