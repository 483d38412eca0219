import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch, numpy as np
from pn2_b200 import pointnet_util as U, scenes, _lib
from pn2_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
torch.manual_seed(0)

def run_sa(N, M, K, D, mlp, radius, B=2):
    pts = torch.from_numpy(scenes.scannet_batch(3, B, N)).to(dev)
    xyz = pts[:, :, :3].contiguous()
    feat = torch.randn(B, N, D, device=dev) if D else None
    layers = []
    c = 3 + D
    for o in mlp:
        layers.append((torch.randn(o, c, device=dev) / c ** 0.5, torch.randn(o, device=dev) * 0.1, True))
        c = o
    f = U.FoldedMlp(layers)
    _, new_xyz = U.fps_gather_cl(xyz, M)
    idx = pu.ball_query(radius, K, xyz, new_xyz)
    U.set_mlp_precision("fp32"); a = U.sa_mlp_max_cl(xyz, feat, new_xyz, idx, _lib.ORDER_XYZ_FIRST, f)
    U.set_mlp_precision("bf16"); b = U.sa_mlp_max_cl(xyz, feat, new_xyz, idx, _lib.ORDER_XYZ_FIRST, f)
    torch.cuda.synchronize()
    err = (a - b).abs().max().item(); sc = a.abs().max().item()
    print("SA N=%d M=%d K=%d D=%d mlp=%s bf16_ok=%s: err %.3e scale %.3e rel %.3e" % (N, M, K, D, mlp, f.bf16_ok(), err, sc, err / sc), flush=True)

def run_fp(n, m, D1, D2, mlp, relu_last=True, B=2):
    pts = torch.from_numpy(scenes.scannet_batch(5, B, n)).to(dev)
    xyz1 = pts[:, :, :3].contiguous(); xyz2 = xyz1[:, :m].contiguous()
    f1 = torch.randn(B, n, D1, device=dev) if D1 else None
    f2 = torch.randn(B, m, D2, device=dev)
    layers = []; c = D1 + D2
    for i, o in enumerate(mlp):
        layers.append((torch.randn(o, c, device=dev) / c ** 0.5, torch.randn(o, device=dev) * 0.1, relu_last or i < len(mlp) - 1)); c = o
    f = U.FoldedMlp(layers)
    idx, w = U.three_nn_weights_cl(xyz1, xyz2)
    U.set_mlp_precision("fp32"); a = U.fp_mlp_cl(f1, f2, idx, w, f, n)
    U.set_mlp_precision("bf16"); b = U.fp_mlp_cl(f1, f2, idx, w, f, n)
    torch.cuda.synchronize()
    err = (a - b).abs().max().item(); sc = a.abs().max().item()
    print("FP n=%d m=%d D1=%d D2=%d mlp=%s bf16_ok=%s: err %.3e scale %.3e rel %.3e" % (n, m, D1, D2, mlp, f.bf16_ok(), err, sc, err / sc), flush=True)

run_sa(1024, 64, 32, 0, [32], 0.3)
run_sa(1024, 64, 32, 3, [32, 32, 64], 0.3)
run_sa(2048, 128, 32, 64, [64, 64, 128], 0.3)
run_sa(512, 32, 32, 256, [256, 256, 512], 0.8)
run_sa(1024, 64, 16, 5, [16, 48], 0.3)
run_sa(1024, 64, 64, 5, [16, 48], 0.5)
run_sa(1024, 64, 128, 5, [16, 48], 0.9)
run_fp(1024, 128, 3, 128, [128, 128, 128, 128, 21], relu_last=False)
run_fp(512, 64, 128, 256, [256, 256])
run_fp(300, 1, 4, 16, [32])
run_fp(700, 100, 0, 128, [128, 64])
