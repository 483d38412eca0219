"""ncu target: the fused lifting at the config-3 shapes (32 scenes x 8192 points, V views of 128 x 32 x 41 maps).
    ncu --set full --clock-control none --import-source on -k regex:lift_ -s <skip> -c <n> -o prof python scripts/prof_lift.py [V]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes
from pn2_b200.projection import lift_views
dev = torch.device("cuda:0")
V = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, N, C = 32, 8192, 128
pts = scenes.scannet_batch(7000, B, N)[:, :, :3].astype(np.float32)
mv = [scenes.multiview_inputs(7000 + b, pts[b], V, C) for b in range(B)]
args = (torch.from_numpy(pts).to(dev), torch.from_numpy(np.stack([m[0] for m in mv])).to(dev),
        torch.from_numpy(np.stack([m[1] for m in mv])).to(dev), torch.from_numpy(np.stack([m[2] for m in mv]).astype(np.float32)).to(dev),
        scenes.SCANNET_INTRINSIC, 0.1, 4.0, scenes.SCANNET_IMAGE_DIMS, 0.05)
for _ in range(3):
    out = lift_views(*args, reduce="max")
torch.cuda.synchronize()
print("ok", out.shape)
