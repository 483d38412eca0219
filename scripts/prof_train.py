import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch, torch.nn.functional as F
from pn2_b200 import scenes
from pn2_b200.models import PointNet2Multiview2Msg
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, C = 8, 8192, 21
net = PointNet2Multiview2Msg(C).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
pts = torch.from_numpy(scenes.scannet_batch(77, B, N)).to(dev)
xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, N, device=dev)
target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1
def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(net(xyz, img).reshape(-1, C), target.reshape(-1), ignore_index=0)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
