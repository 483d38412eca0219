"""MSG semseg train step (BASELINE config 2) on one GPU: fused training kernels (csrc/train_mlp.cu) vs the torch.nn
conv / BatchNorm / ReLU composition, same geometry kernels.  CUDA events, median of the steps after two warm-ups."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch, torch.nn.functional as F
from pn2_b200 import scenes, pointnet_util
from pn2_b200.models import PointNet2Multiview2Msg, PointNet2SemSeg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name in ("msg", "ssg"):
    for fused in (True, False):
        pointnet_util.set_fused_training(fused)
        torch.manual_seed(0)
        net = (PointNet2Multiview2Msg(21) if name == "msg" else PointNet2SemSeg(21)).to(dev).train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
        pts = torch.from_numpy(scenes.scannet_batch(77, B, 8192)).to(dev)
        xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
        second = torch.randn(B, 128, 8192, device=dev) if name == "msg" else pts[:, :, 3:].permute(0, 2, 1).contiguous()
        target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1
        evs, losses = [], []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(net(xyz, second).reshape(-1, 21), target.reshape(-1), ignore_index=0)
            loss.backward()
            opt.step()
            b.record()
            evs.append((a, b)); losses.append(loss.detach())
        torch.cuda.synchronize()
        ms = float(np.median([x.elapsed_time(y) for x, y in evs[2:]]))
        print(json.dumps({"model": name, "fused_training": fused, "batch": B, "ms_per_step": round(ms, 3), "scenes_per_s": round(B / ms * 1e3, 1),
                          "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
        del net, opt
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
