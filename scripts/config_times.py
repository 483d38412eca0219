"""Forward time of the other BASELINE configs on one GPU (inference, fused path, CUDA events, 3 warm-ups):
config 2 (MSG stack forward; the train step is examples/train_msg_semseg_ddp.py), config 3 (multi-view: lifting + the
PointNet2Multiview2 point branch), config 4 (nuScenes backbone, 16384-point and ~35k-point sweeps, batch 16)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import scenes
from pn2_b200.models import GraphedForward, PipelinedForward, PointNet2Backbone, PointNet2Multiview2, PointNet2Multiview2Msg

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timeit(fn, iters=10):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


B, N = 32, 8192
pts = scenes.scannet_batch(0, B, N).astype(np.float32)
xyz = torch.from_numpy(pts[:, :, :3]).to(dev).permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, N, device=dev)
msg = PointNet2Multiview2Msg(21).eval().to(dev)
ms = timeit(lambda: msg(xyz, img))
print(json.dumps({"config": 2, "what": "PointNet2Multiview2Msg point branch forward (MSG SA x6, FP x4, head), batch 32 x 8192, eager single stream",
                  "ms": ms, "scenes_per_s": B / ms * 1e3}))


def graphed_and_pipelined(model, a, b, tag, cfg):
    g = GraphedForward(model, a, b)
    ms = timeit(lambda: g.run(a, b))
    print(json.dumps({"config": cfg, "what": tag + ", CUDA graph replay, one batch at a time", "ms": ms, "scenes_per_s": B / ms * 1e3}))
    pipe = PipelinedForward(model, a, b, depth=6)
    for _ in range(12):
        pipe.submit(a, b)
    pipe.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60):
        pipe.submit(a, b)
    pipe.join()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 60
    print(json.dumps({"config": cfg, "what": tag + ", 6 graph instances in flight", "ms": ms, "scenes_per_s": B / ms * 1e3}))


graphed_and_pipelined(msg, xyz, img, "PointNet2Multiview2Msg point branch forward", 2)
V = 3
mv = [scenes.multiview_inputs(7000 + b, pts[b, :, :3], V, 128) for b in range(B)]
feats = torch.from_numpy(np.stack([m[0] for m in mv])).to(dev)
depth = torch.from_numpy(np.stack([m[1] for m in mv])).to(dev)
poses = torch.from_numpy(np.stack([m[2] for m in mv]).astype(np.float32)).to(dev)
ssg = PointNet2Multiview2(21).eval().to(dev)
ms = timeit(lambda: ssg.forward_views(xyz, feats, depth, poses, scenes.SCANNET_INTRINSIC, 0.1, 4.0, scenes.SCANNET_IMAGE_DIMS, 0.05))
print(json.dumps({"config": 3, "what": "PointNet2Multiview2.forward_views: lifting of 3 views (128 x 32 x 41 maps, ENet output as input) + point branch, "
                  "batch 32 x 8192, eager single stream", "ms": ms, "scenes_per_s": B / ms * 1e3}))
graphed_and_pipelined(ssg, xyz, img, "PointNet2Multiview2 point branch forward (lifted features resident)", 3)
bb = PointNet2Backbone().eval().to(dev)
for n in (16384, 34720):
    sw = [scenes.lidar_sweep(50 + i, n) for i in range(16)]  # (xyz (n, 3), feat (n, 2)) per sweep
    x3 = torch.from_numpy(np.stack([s[0] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
    f2 = torch.from_numpy(np.stack([s[1] for s in sw]).astype(np.float32)).to(dev).permute(0, 2, 1).contiguous()
    ms = timeit(lambda: bb(x3, f2), iters=5)
    print(json.dumps({"config": 4, "what": "nuScenes backbone (model/pointmaskrcnn.py:8-32), batch 16 x %d points, eager" % n, "ms": ms,
                      "sweeps_per_s": 16 / ms * 1e3}))
