"""MSG semseg train step: eager launches vs ONE CUDA graph of forward + backward + Adam (developer timing)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch, torch.nn.functional as F
from pn2_b200 import scenes
from pn2_b200.models import PointNet2Multiview2Msg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4


def make():
    torch.manual_seed(0)
    net = PointNet2Multiview2Msg(21).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    return net, opt


pts = torch.from_numpy(scenes.scannet_batch(77, B, 8192)).to(dev)
xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
img = torch.randn(B, 128, 8192, device=dev)
target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1


def step(net, opt):
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(net(xyz, img).reshape(-1, 21), target.reshape(-1), ignore_index=0)
    loss.backward()
    opt.step()
    return loss


def timed(fn, n=10):
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([x.elapsed_time(y) for x, y in evs]))


net, opt = make()
losses_e = [step(net, opt).detach() for _ in range(3)]
ms_e = timed(lambda: step(net, opt))
losses_e = [float(v) for v in losses_e]
net, opt = make()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        l = step(net, opt)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
net2, opt2 = make()  # fresh weights: replay the first three steps from the same start as the eager run
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(net2, opt2)  # one eager step allocates grads / Adam state before capture
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
opt2.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    static_loss = step(net2, opt2)
losses_g = [losses_e[0]]
for _ in range(2):
    g.replay()
    losses_g.append(float(static_loss))
ms_g = timed(g.replay)
print(json.dumps({"batch": B, "eager_ms": round(ms_e, 3), "graph_ms": round(ms_g, 3), "eager_losses": losses_e, "graph_losses(from step 2)": losses_g}))
