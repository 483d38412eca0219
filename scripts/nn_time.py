"""3-NN search time on a prebuilt cell list (fp1 shape: 8192 queries against the 1024 sampled points, 32 scenes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes
from pn2_b200.pointnet_util import SpatialGrid, fps_gather_cl
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
xyz = pts[:, :, :3].contiguous()
for (n, m) in ((8192, 1024), (1024, 256), (8192, 2048), (8192, 4096)):
    fine = xyz[:, :n].contiguous()
    _, coarse = fps_gather_cl(fine, m)
    g_fine, g = SpatialGrid(fine, 0.0), SpatialGrid(coarse, 0.0)
    for order in (None, g_fine.order):
        ts = []
        for i in range(12):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.three_nn(fine, query_order=order); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("three_nn n=%d m=%d %s: %.1f us" % (n, m, "sorted queries" if order is not None else "index order", 1e3 * float(np.median(ts[2:]))))
