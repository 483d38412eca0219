"""Throughput of the fused lifting (pn2_lift_views) on BASELINE config 3 shapes: batch x 8192 points, V views of 128 x 32 x 41
ENet-sized feature maps.  Algorithmic bytes per scene: 12N + V(4HW + 64) + 4C*min(VHW, VN) + 4CN."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes
from pn2_b200.projection import lift_views

dev = torch.device("cuda:0")
B, N, C, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 8192, 128, 32, 41
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for V in (3, 5):
    pts = scenes.scannet_batch(7000, B, N)[:, :, :3].astype(np.float32)
    mv = [scenes.multiview_inputs(7000 + b, pts[b], V, C) for b in range(B)]
    feats = torch.from_numpy(np.stack([m[0] for m in mv])).to(dev)
    depth = torch.from_numpy(np.stack([m[1] for m in mv])).to(dev)
    poses = torch.from_numpy(np.stack([m[2] for m in mv]).astype(np.float32)).to(dev)
    p = torch.from_numpy(pts).to(dev)
    args = (p, feats, depth, poses, scenes.SCANNET_INTRINSIC, 0.1, 4.0, scenes.SCANNET_IMAGE_DIMS, 0.05)
    for red in ("max", "first"):
        for _ in range(3):
            out = lift_views(*args, reduce=red)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = lift_views(*args, reduce=red); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        byts = B * (12 * N + V * (4 * H * W + 64) + 4 * C * min(V * H * W, V * N) + 4 * C * N)
        lifted = float((out != 0).any(dim=1).float().mean())
        print(json.dumps({"op": "lift_views", "B": B, "V": V, "reduce": red, "ms": ms, "scenes_per_s": B / ms * 1e3, "gbs": byts / ms / 1e6,
                          "frac_hbm": byts / ms / 1e6 / PEAK, "points_lifted": lifted,
                          "note": "includes the host-side pose inverse / frustum setup (torch) of the wrapper"}))

# kernel-only time of the last configuration (CUDA events around the ABI call)
from pn2_b200 import _lib
_lib.PROFILE = []
for _ in range(5):
    lift_views(*args, reduce="max")
torch.cuda.synchronize()
print("kernel only (max, V=5): %.3f ms" % float(np.median([a.elapsed_time(b) for _, a, b in _lib.PROFILE])))
_lib.PROFILE = None
