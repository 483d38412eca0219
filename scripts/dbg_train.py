import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch, torch.nn as nn, torch.nn.functional as F
from pn2_b200 import train_mlp, _lib
from pn2_b200._lib import ptr
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
R, cin, cout = 1000, 7, 16
x = torch.randn(R, cin, device=dev)
W = torch.randn(cout, cin, device=dev) * 0.3; b = torch.randn(cout, device=dev) * 0.1
gamma = torch.rand(cout, device=dev) + 0.5; beta = torch.randn(cout, device=dev) * 0.2
g = torch.randn(R, cout, device=dev)
# torch reference with autograd on z
xr = x.clone().requires_grad_(True); Wr = W.clone().requires_grad_(True)
z = xr @ Wr.t() + b; z.retain_grad()
mean = z.mean(0); var = z.var(0, unbiased=False); rstd = (var + 1e-5).rsqrt()
y = (z - mean) * rstd * gamma + beta
a = torch.relu(y)
(a * g).sum().backward()
# ours
zz = torch.empty(R, cout, device=dev); stats = torch.zeros(2, cout, dtype=torch.float64, device=dev)
s = _lib.stream_ptr(dev)
_lib.call("pn2_train_linear_fwd", R, cin, cout, ptr(x), None, None, ptr(W), ptr(b), ptr(zz), ptr(stats), s)
print("z err", float((zz - z).abs().max()), "mean err", float((stats[0] / R - mean.double()).abs().max()))
m64 = stats[0] / R; v64 = stats[1] / R - m64 * m64; r64 = torch.rsqrt(v64 + 1e-5)
scale = (gamma.double() * r64).float(); shift = (beta.double() - m64 * gamma.double() * r64).float()
sums = torch.zeros(2, cout, dtype=torch.float64, device=dev)
_lib.call("pn2_train_bn_bwd_reduce", R, cout, ptr(g), ptr(zz), ptr(scale), ptr(shift), ptr(sums), s)
dy = g * (y > 0)
print("S1 err", float((sums[0] - dy.double().sum(0)).abs().max()), "S2 err", float((sums[1] - (dy * z).double().sum(0)).abs().max()))
S1, S2 = sums
t = S2 - m64 * S1
ca = gamma.double() * r64; cc = -gamma.double() * r64 ** 3 * t / R; cb = -ca * S1 / R - cc * m64
dz_mine = ca.float() * dy + cb.float() + cc.float() * z
print("dz formula err", float((dz_mine - z.grad).abs().max()), float(z.grad.abs().max()))
gin = torch.empty(R, cin, device=dev); dW = torch.zeros(cout, cin, device=dev)
_lib.call("pn2_train_linear_bwd", R, cin, cout, ptr(x), None, None, ptr(W.t().contiguous()), ptr(g), ptr(zz), ptr(scale), ptr(shift),
          ptr(ca.float().contiguous()), ptr(cb.float().contiguous()), ptr(cc.float().contiguous()), ptr(gin), ptr(dW), s)
torch.cuda.synchronize()
print("g_in err", float((gin - xr.grad).abs().max()), float(xr.grad.abs().max()), "vs dz@W", float((gin - dz_mine.detach() @ W).abs().max()))
print("dW err", float((dW - Wr.grad).abs().max()), float(Wr.grad.abs().max()))
