"""Solo duration of every ABI call of the nuScenes backbone forward (config 4), batch 16."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes, _lib
from pn2_b200.models import PointNet2Backbone
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = PointNet2Backbone().eval().to(dev)
for n in (int(a) for a in (sys.argv[1:] or ["34720"])):
    sw = [scenes.lidar_sweep(50 + i, n) for i in range(16)]
    x3 = torch.from_numpy(np.stack([s[0] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
    f2 = torch.from_numpy(np.stack([s[1] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
    with torch.no_grad():
        for _ in range(2):
            model(x3, f2)
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for i in range(3):
            _lib.PROFILE = []
            model(x3, f2)
            torch.cuda.synchronize()
            for j, (name, a, b) in enumerate(_lib.PROFILE):
                agg.setdefault((j, name), []).append(a.elapsed_time(b))
        _lib.PROFILE = None
    print("== n =", n)
    tot = 0
    for (j, name), v in agg.items():
        ms = sorted(v)[len(v) // 2]
        tot += ms
        print("%2d %-28s %8.1f us" % (j, name, ms * 1e3))
    print("sum %.1f us" % (tot * 1e3))
