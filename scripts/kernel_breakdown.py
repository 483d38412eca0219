"""Warm, solo duration of every ABI call of a fused forward (single stream, CUDA events around each call; pn2_b200._lib.PROFILE).

    python scripts/kernel_breakdown.py ssg [batch]          SSG semseg, 8192 points (BASELINE config 1)
    python scripts/kernel_breakdown.py mv                   point branches of the two multi-view stacks (configs 2 / 3)
    python scripts/kernel_breakdown.py bb [n ...]           nuScenes backbone, batch 16 (config 4; default 34720 points)
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pn2_b200 import _lib, pointnet_util, scenes  # noqa: E402
from pn2_b200.models import PointNet2Backbone, PointNet2Multiview2, PointNet2Multiview2Msg, PointNet2SemSeg  # noqa: E402

dev = torch.device("cuda:0")
pointnet_util.set_mlp_precision(os.environ.get("PRECISION", "bf16"))


def breakdown(title, model, batches, reps):
    model.single_stream = True
    with torch.no_grad():
        for i in range(3):
            model(*batches[i % len(batches)])
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for i in range(reps):
            _lib.PROFILE = []
            model(*batches[i % len(batches)])
            torch.cuda.synchronize()
            for j, (name, a, b) in enumerate(_lib.PROFILE):
                agg.setdefault((j, name), []).append(a.elapsed_time(b))
        _lib.PROFILE = None
    print("==", title)
    tot = 0.0
    for (j, name), v in agg.items():
        ms = sorted(v)[len(v) // 2]
        tot += ms
        print("%2d %-28s %8.1f us" % (j, name, ms * 1e3))
    print("sum %.1f us" % (tot * 1e3))


what = sys.argv[1] if len(sys.argv) > 1 else "ssg"
torch.manual_seed(0)
if what == "ssg":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    xs = [torch.from_numpy(scenes.scannet_batch(100 * i, B, 8192)).to(dev).permute(0, 2, 1).contiguous() for i in range(4)]
    breakdown("PointNet2SemSeg, batch %d" % B, PointNet2SemSeg(21).eval().to(dev), [(x[:, :3], x[:, 3:]) for x in xs], 10)
elif what == "mv":
    xyz = torch.from_numpy(scenes.scannet_batch(0, 32, 8192)[:, :, :3]).to(dev).permute(0, 2, 1).contiguous()
    img = torch.randn(32, 128, 8192, device=dev)
    for cls in (PointNet2Multiview2, PointNet2Multiview2Msg):
        breakdown(cls.__name__, cls(21).eval().to(dev), [(xyz, img)], 6)
else:
    model = PointNet2Backbone().eval().to(dev)
    for n in (int(a) for a in (sys.argv[2:] or ["34720"])):
        sw = [scenes.lidar_sweep(50 + i, n) for i in range(16)]
        x3 = torch.from_numpy(np.stack([s[0] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
        f2 = torch.from_numpy(np.stack([s[1] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
        breakdown("PointNet2Backbone, 16 sweeps x %d points" % n, model, [(x3, f2)], 6)
