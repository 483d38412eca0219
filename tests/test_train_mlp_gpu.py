"""The fused training-mode shared MLP (csrc/train_mlp.cu, pn2_b200/train_mlp.py) against torch's own conv -> BatchNorm
(batch statistics) -> ReLU chain under autograd on the same device (TF32 off): outputs, every gradient and the running
statistics.  Row counts that are not multiples of the 128-row tile and widths that are not multiples of 4 are covered."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from pn2_b200 import train_mlp

pytestmark = pytest.mark.gpu


def stack(widths, seed):
    torch.manual_seed(seed)
    convs, bns = nn.ModuleList(), nn.ModuleList()
    for cin, cout in zip(widths[:-1], widths[1:]):
        convs.append(nn.Conv1d(cin, cout, 1))
        bn = nn.BatchNorm1d(cout)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
            bn.running_mean.normal_(0, 0.1)
            bn.running_var.uniform_(0.5, 1.5)
        bns.append(bn)
    return convs, bns


@pytest.mark.parametrize("R,widths", [(1000, [7, 16, 32]), (4096 + 13, [131, 32, 32, 64]), (2048, [6, 32, 32, 64]),
                                      (777, [515, 128, 196, 256]), (300, [1536, 512, 512]), (129, [3, 5])])
def test_fused_training_chain_matches_torch_autograd(cuda, R, widths):
    convs, bns = stack(widths, R)
    convs, bns = convs.to(cuda).train(), bns.to(cuda).train()
    ref_convs, ref_bns = copy.deepcopy(convs), copy.deepcopy(bns)
    g = torch.Generator().manual_seed(R)
    x0 = torch.randn(R, widths[0], generator=g).to(cuda)
    probe = torch.randn(R, widths[-1], generator=g).to(cuda)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        xa = x0.clone().requires_grad_(True)
        out = train_mlp.fused_mlp_train(xa, convs, bns)
        (out * probe).sum().backward()
        xb = x0.clone().requires_grad_(True)
        y = xb.t().unsqueeze(0)                      # (1, C, R)
        for conv, bn in zip(ref_convs, ref_bns):
            y = F.relu(bn(conv(y)))
        ref = y.squeeze(0).t()
        (ref * probe).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32

    def close(a, b, rel, what):
        scale = float(b.detach().abs().max().clamp_min(1e-12))
        err = float((a.detach() - b.detach()).abs().max())
        assert err <= rel * scale, "%s: max abs err %.3e vs %.1e * %.3e" % (what, err, rel, scale)

    close(out, ref, 2e-5, "output")
    close(xa.grad, xb.grad, 2e-4, "input gradient")
    wscale = max(float(c.weight.grad.abs().max()) for c in ref_convs)
    for l, (c, rc, b, rb) in enumerate(zip(convs, ref_convs, bns, ref_bns)):
        close(c.weight.grad, rc.weight.grad, 2e-4, "dW%d" % l)
        # the true bias gradient is exactly zero in front of a batch-statistics BatchNorm (torch holds rounding noise)
        assert float(c.bias.grad.abs().max()) == 0.0 and float(rc.bias.grad.abs().max()) <= 1e-3 * wscale * R ** 0.5
        close(b.weight.grad, rb.weight.grad, 2e-4, "dgamma%d" % l)
        close(b.bias.grad, rb.bias.grad, 2e-4, "dbeta%d" % l)
        close(b.running_mean, rb.running_mean, 1e-5, "running_mean%d" % l)
        close(b.running_var, rb.running_var, 1e-5, "running_var%d" % l)
        assert int(b.num_batches_tracked) == int(rb.num_batches_tracked) == 1
