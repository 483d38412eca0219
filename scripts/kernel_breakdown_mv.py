"""Solo duration of every ABI call of the multi-view point branch (single stream, CUDA events around each call)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import scenes, _lib
from pn2_b200.models import PointNet2Multiview2, PointNet2Multiview2Msg
dev = torch.device("cuda:0")
B = 32
torch.manual_seed(0)
for cls in (PointNet2Multiview2, PointNet2Multiview2Msg):
    model = cls(21).eval().to(dev)
    xyz = torch.from_numpy(scenes.scannet_batch(0, B, 8192)[:, :, :3]).to(dev).permute(0, 2, 1).contiguous()
    img = torch.randn(B, 128, 8192, device=dev)
    with torch.no_grad():
        for _ in range(3):
            model(xyz, img)
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for i in range(6):
            _lib.PROFILE = []
            model(xyz, img)
            torch.cuda.synchronize()
            for j, (name, a, b) in enumerate(_lib.PROFILE):
                agg.setdefault((j, name), []).append(a.elapsed_time(b))
        _lib.PROFILE = None
    print("==", cls.__name__)
    tot = 0
    for (j, name), v in agg.items():
        ms = sorted(v)[len(v) // 2]
        tot += ms
        print("%2d %-28s %8.1f us" % (j, name, ms * 1e3))
    print("sum %.1f us" % (tot * 1e3))
