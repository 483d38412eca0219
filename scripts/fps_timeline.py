import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import pointnet_util as U, scenes, _lib
dev = torch.device("cuda:0")
pts = torch.from_numpy(scenes.scannet_batch(0, 32, 8192)).to(dev)
xyz = pts[:, :, :3].contiguous()
lib = _lib.load()
lib.pn2_debug_set_fps_stamps.argtypes = [ctypes.c_void_p]
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib.pn2_debug_set_fps_mode(2)
U.fps_gather_cl(xyz, 1024); torch.cuda.synchronize()
lib.pn2_debug_set_fps_stamps(buf.data_ptr())
U.fps_gather_cl(xyz, 1024); torch.cuda.synchronize()
t = buf.cpu().view(-1, 4)[:10].tolist()
for r in t:
    print("compute+redux %5d  send..wait-done %5d  decode %5d  | round-to-round" % (r[1]-r[0], r[2]-r[1], r[3]-r[2]))
print("round starts deltas:", [t[i+1][0]-t[i][0] for i in range(9)])
