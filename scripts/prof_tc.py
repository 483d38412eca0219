"""ncu target: one fp1+head-shaped and one sa1-shaped launch of the tensor-core MLP kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import pointnet_util as U, scenes, _lib
from pn2_b200 import pointnet2_utils as pu
from pn2_b200.models import PointNet2SemSeg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
U.set_mlp_precision(os.environ.get("PRECISION", "bf16"))  # fp32: the FFMA kernels of row_mlp.cu
torch.manual_seed(0)
model = PointNet2SemSeg(21).eval().to(dev)
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
x6 = pts.permute(0, 2, 1).contiguous()
with torch.no_grad():
    for _ in range(3):
        y = model(x6[:, :3], x6[:, 3:])
torch.cuda.synchronize()
print("ok", y.shape)
