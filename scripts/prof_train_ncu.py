"""ncu target: two eager MSG train steps (8 scenes); profile the training GEMM kernels of the second one, e.g.
ncu --set full -k regex:train_linear --launch-skip <launches of step 1> -c N python scripts/prof_train_ncu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch, torch.nn.functional as F
from pn2_b200 import scenes
from pn2_b200.models import PointNet2Multiview2Msg
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, C = int(os.environ.get("B", "8")), 8192, 21
net = PointNet2Multiview2Msg(C).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
pts = torch.from_numpy(scenes.scannet_batch(77, B, N)).to(dev)
xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, N, device=dev)
target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1
for _ in range(int(os.environ.get("STEPS", "2"))):
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(net(xyz, img).reshape(-1, C), target.reshape(-1), ignore_index=0)
    loss.backward(); opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
