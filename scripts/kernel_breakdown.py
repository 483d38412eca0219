"""Warm, solo duration of every kernel of the fused forward (single stream, CUDA events around each ABI call)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import scenes, _lib
from pn2_b200.models import PointNet2SemSeg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
model = PointNet2SemSeg(21).eval().to(dev)
model.single_stream = True
xs = [torch.from_numpy(scenes.scannet_batch(100 * i, B, 8192)).to(dev).permute(0, 2, 1).contiguous() for i in range(4)]
with torch.no_grad():
    for i in range(3):
        model(xs[i % 4][:, :3], xs[i % 4][:, 3:])
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    R = 10
    for i in range(R):
        _lib.PROFILE = []
        model(xs[i % 4][:, :3], xs[i % 4][:, 3:])
        torch.cuda.synchronize()
        for j, (name, a, b) in enumerate(_lib.PROFILE):
            agg.setdefault((j, name), []).append(a.elapsed_time(b))
    _lib.PROFILE = None
tot = 0
for (j, name), v in agg.items():
    ms = sorted(v)[len(v) // 2]
    tot += ms
    print("%2d %-28s %8.1f us" % (j, name, ms * 1e3))
print("sum %.1f us" % (tot * 1e3))
