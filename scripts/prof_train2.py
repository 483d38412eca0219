"""Kernel time by name of one MSG train step (torch.profiler, CUDA activities), fused vs unfused training path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from pn2_b200 import scenes, pointnet_util
from pn2_b200.models import PointNet2Multiview2Msg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fused = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
pointnet_util.set_fused_training(fused)
torch.manual_seed(0)
net = PointNet2Multiview2Msg(21).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
pts = torch.from_numpy(scenes.scannet_batch(77, B, 8192)).to(dev)
xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, 8192, device=dev)
target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1
def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(net(xyz, img).reshape(-1, 21), target.reshape(-1), ignore_index=0)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
