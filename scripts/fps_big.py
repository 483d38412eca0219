import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes
from pn2_b200.pointnet_util import fps_gather_cl
dev = torch.device("cuda:0")
from pn2_b200 import _lib
REF = {}
for mode in (0,):
  _lib.load().pn2_debug_set_fps_mode(mode)
  print("mode", mode)
  for B, n, m in ((16, 34720, 4096), (16, 16384, 4096)):
      x = torch.from_numpy(np.stack([scenes.lidar_sweep(50 + i, n)[0] for i in range(B)]).astype(np.float32)).to(dev)
      for _ in range(2): fps_gather_cl(x, m)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(3): fps_gather_cl(x, m)
      b.record(); torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 3
      print("B=%d n=%d m=%d: %.3f ms, %.3f us/round" % (B, n, m, ms, ms * 1e3 / (m - 1)))
      idx = fps_gather_cl(x, m)[0]
      ref = REF.setdefault((B, n, m), idx.clone())
      assert torch.equal(ref, idx), "indices differ between kernel variants"
