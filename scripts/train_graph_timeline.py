"""Kernel timeline of one replayed MSG train step (GraphedTrainStep, 4 scenes): torch.profiler over two replays; prints
the wall time of a replay, the kernel time per stream, the time with no kernel running and the longest kernels."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from pn2_b200 import scenes
from pn2_b200.models import GraphedTrainStep, PointNet2Multiview2Msg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
net = PointNet2Multiview2Msg(21).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
pts = torch.from_numpy(scenes.scannet_batch(77, B, 8192)).to(dev)
xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, 8192, device=dev)
target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1
loss_fn = lambda lg, t: F.cross_entropy(lg.reshape(-1, 21), t.reshape(-1), ignore_index=0)
stepper = GraphedTrainStep(net, opt, loss_fn, xyz, img, target)
for _ in range(3):
    stepper.step(xyz, img, target)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    stepper.step(xyz, img, target)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
print("kernels %d, wall %.2f ms, summed kernel time %.2f ms" % (len(ev), (t1 - t0) / 1e3, sum(e.time_range.end - e.time_range.start for e in ev) / 1e3))
# time covered by at least one kernel
cover, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None: cover += cur_e - cur_s
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
cover += cur_e - cur_s
print("time with >= 1 kernel running %.2f ms, idle gaps %.2f ms" % (cover / 1e3, (t1 - t0 - cover) / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:70]]; a[0] += 1; a[1] += (e.time_range.end - e.time_range.start)
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-72s n=%4d %8.1f us  each %6.1f" % (name, n, us, us / n))
# quarter-by-quarter concurrency
edges = np.linspace(t0, t1, 9)
for a, b in zip(edges[:-1], edges[1:]):
    busy = sum(max(0.0, min(e.time_range.end, b) - max(e.time_range.start, a)) for e in ev)
    print("window %.2f-%.2f ms: mean kernels in flight %.2f" % ((a - t0) / 1e3, (b - t0) / 1e3, busy / (b - a)))
