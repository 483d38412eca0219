"""ncu target for the HBM-bound operators at the C1 shapes: three_interpolate (tiled kernel and the lane-along-channel
kernel), group_points, gather_points and the fused lifting.  Only the last launch of each is inside the profiler range.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o prof_ops python scripts/prof_ops.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import _lib, scenes
from pn2_b200 import pointnet2_utils as pu
from pn2_b200.projection import lift_views
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
B, C, m, n = 32, 128, 1024, 8192
f = torch.randn(B, C, m, device=dev)
idx = torch.randint(0, m, (B, n, 3), device=dev, dtype=torch.int32, generator=g)
w = torch.rand(B, n, 3, device=dev); w = (w / w.sum(-1, keepdim=True)).contiguous()
f64 = torch.randn(B, 64, n, device=dev)
gidx = torch.randint(0, n, (B, 1024, 32), device=dev, dtype=torch.int32, generator=g)
f128 = torch.randn(64, 128, 16384, device=dev)
sidx = torch.randint(0, 16384, (64, 4096), device=dev, dtype=torch.int32, generator=g)
V, CH = 3, 128
pts = scenes.scannet_batch(7000, B, n)[:, :, :3].astype(np.float32)
mv = [scenes.multiview_inputs(7000 + b, pts[b], V, CH) for b in range(B)]
largs = (torch.from_numpy(pts).to(dev), torch.from_numpy(np.stack([x[0] for x in mv])).to(dev),
         torch.from_numpy(np.stack([x[1] for x in mv])).to(dev), torch.from_numpy(np.stack([x[2] for x in mv]).astype(np.float32)).to(dev),
         scenes.SCANNET_INTRINSIC, 0.1, 4.0, scenes.SCANNET_IMAGE_DIMS, 0.05)


ONLY_INTERP = len(sys.argv) > 1 and sys.argv[1] == "interp"


def ops():
    for mode in (1, 32):
        lib.pn2_debug_set_interp_mode(mode)
        pu.three_interpolate(f, idx, w)
    lib.pn2_debug_set_interp_mode(0)
    if ONLY_INTERP:
        return
    pu.grouping_operation(f64, gidx)
    pu.gather_operation(f128, sidx)
    lift_views(*largs, reduce="max")


for _ in range(3):
    ops()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
