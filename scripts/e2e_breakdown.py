import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from diffdrr_b200 import DRR, synthetic
dev = torch.device("cuda:0"); D, H, B = 512, 256, 16
vol = torch.rand(D, D, D, device=dev)
subj = synthetic.make_subject(torch.zeros(1, 1, 1, 1)); subj.volume.affine = synthetic.make_affine(D)
drr = DRR(subj, **synthetic.detector_kwargs(H)).to(dev); drr.density = vol
rot_h, xyz_h = synthetic.make_poses(B, seed=0); rot_h, xyz_h = rot_h.pin_memory(), xyz_h.pin_memory()
w = torch.rand(B, 1, H, H, device=dev)
img_h = torch.empty(B, 1, H, H).pin_memory()
def step():
    rot = rot_h.to(dev, non_blocking=True).requires_grad_(True); xyz = xyz_h.to(dev, non_blocking=True).requires_grad_(True)
    img = drr(rot, xyz, parameterization="euler_angles", convention="ZXY")
    loss = (img * w).sum(); loss.backward()
    img_h.copy_(img.detach(), non_blocking=True); torch.cuda.current_stream().synchronize()
for _ in range(3): step()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t0) / 20 * 1e3)
