import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import pointnet_util as U, scenes, _lib
from pn2_b200.models import PointNet2SemSeg
dev = torch.device("cuda:0")
B = 32
torch.manual_seed(0)
model = PointNet2SemSeg(21).eval().to(dev)
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
x6 = pts.permute(0, 2, 1).contiguous()
lib = _lib.load()
lib.pn2_debug_set_tc_timestamps.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    for _ in range(2): model(x6[:, :3], x6[:, 3:])
    xyz_cl, feat_cl = U.to_channel_last(x6[:, :3]), U.to_channel_last(x6[:, 3:])
    l1_xyz, l1 = model.sa1.forward_cl(xyz_cl, feat_cl)
    l2_xyz, l2 = model.sa2.forward_cl(l1_xyz, l1)
    nnw = U.three_nn_weights_cl(xyz_cl, l1_xyz)
    l1n = torch.randn_like(l1[:, :, :1]).expand(-1, -1, 128).contiguous()
    g0 = U.SpatialGrid(xyz_cl, 0.101)
    for name, fn in [("fp1+head sorted", lambda: model.fp1.forward_cl(xyz_cl, l1_xyz, feat_cl, l1n, mlp=model._fp1_with_head(), nn_weights=nnw, row_order=g0.order)),
                     ("fp1+head unsorted", lambda: model.fp1.forward_cl(xyz_cl, l1_xyz, feat_cl, l1n, mlp=model._fp1_with_head(), nn_weights=nnw)),
                     ("sa1", lambda: model.sa1.forward_cl(xyz_cl, feat_cl)),
                     ("sa2", lambda: model.sa2.forward_cl(l1_xyz, l1))]:
        buf = torch.zeros(256, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        lib.pn2_debug_set_tc_timestamps(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        t = buf.cpu().tolist()
        t = [v for v in t if v]
        d = [b - a for a, b in zip(t, t[1:])]
        print(name, "n stamps", len(t), "deltas (cycles):", d[:40])
