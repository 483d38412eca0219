"""Atomic vs deterministic backward (pn2_scatter_rows_det) on the shapes of the SSG network, batch 32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import pointnet2_utils as pu, scenes
from pn2_b200.pointnet_util import fps_gather_cl, three_nn_weights_cl
dev = torch.device("cuda:0")
B = 32
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
xyz = pts[:, :, :3].contiguous()
_, new_xyz = fps_gather_cl(xyz, 1024)
bq = pu.ball_query(0.1, 32, xyz, new_xyz)
i3, w3 = three_nn_weights_cl(xyz, new_xyz)


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


cases = {
    "group_points_grad C=64 N=8192 M=1024 K=32": (lambda f: pu.grouping_operation(f, bq), (B, 64, 8192), (B, 64, 1024, 32)),
    "three_interpolate_grad C=128 n=8192 m=1024": (lambda f: pu.three_interpolate(f, i3, w3), (B, 128, 1024), (B, 128, 8192)),
}
for name, (op, fshape, gshape) in cases.items():
    f = torch.randn(*fshape, device=dev, requires_grad=True)
    go = torch.randn(*gshape, device=dev)
    out = op(f)
    res = {}
    for det in (False, True):
        pu.set_deterministic(det)
        res[det] = timeit(lambda: torch.autograd.grad(out, f, go, retain_graph=True))
    pu.set_deterministic(None)
    print("%-46s atomics %.3f ms   deterministic %.3f ms (inverse index rebuilt every call)" % (name, res[False], res[True]))
