"""Training-mode parity (BASELINE config 2 is a train step): forward with batch-statistics BatchNorm and the backward
kernels (group_points_grad / three_interpolate_grad scatter-adds) inside the real SA -> FP composition, against a
differentiable CPU restatement.  fp32; tolerance 2e-4 relative to the largest gradient entry (atomic accumulation
order differs run to run, as in the reference)."""
import copy

import numpy as np
import pytest
import torch

from oracle import train_ref
from pn2_b200 import scenes
from pn2_b200.pointnet_util import PointNetFeaturePropagation, PointNetSetAbstraction

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_sa_fp_train_step_matches_reference_gradients(cuda):
    torch.manual_seed(0)
    B, N, D = 2, 2048, 6
    sa = PointNetSetAbstraction(256, 0.25, 32, D + 3, [16, 32], False).train()
    fp = PointNetFeaturePropagation(D + 32, [24, 8]).train()
    pts = scenes.scannet_batch(40, B, N)
    xyz = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1).contiguous()
    feat = torch.randn(B, D, N)
    target = torch.randn(B, 8, N)

    def run(sa_m, fp_m, dev, ref):
        x, f = xyz.to(dev), feat.clone().to(dev).requires_grad_(True)
        if ref:
            new_xyz, l1 = train_ref.sa_train_ref(sa_m, x, f)
            out = train_ref.fp_train_ref(fp_m, x, new_xyz, f, l1)
        else:
            new_xyz, l1 = sa_m(x, f)
            out = fp_m(x, new_xyz, f, l1)
        loss = ((out - target.to(dev)) ** 2).mean()
        loss.backward()
        return loss, f.grad, [p.grad for p in list(sa_m.parameters()) + list(fp_m.parameters())]

    sa_g, fp_g = copy.deepcopy(sa).to(cuda), copy.deepcopy(fp).to(cuda)
    loss_ref, fgrad_ref, pgrads_ref = run(sa, fp, "cpu", True)
    # the 1x1 convs of the training path are torch.nn (cuDNN) as in the reference; pin them to true fp32 for the comparison
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        loss, fgrad, pgrads = run(sa_g, fp_g, cuda, False)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    # conv biases feeding a train-mode BatchNorm have an exactly-zero true gradient (both sides hold ~1e-9 noise):
    # errors are therefore measured against the largest parameter gradient, not per tensor
    scale = max(float(r.abs().max()) for r in pgrads_ref)
    errs = [rel_err(fgrad, fgrad_ref)] + [float((g.detach().cpu() - r).abs().max()) / max(float(r.abs().max()), 0.05 * scale)
                                          for g, r in zip(pgrads, pgrads_ref)]
    print("loss", float(loss.detach()), float(loss_ref.detach()), "max rel grad err", max(errs))
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-5 * abs(float(loss_ref.detach()))
    assert max(errs) <= 2e-4, errs
    # the running statistics were updated identically (momentum 0.1, batch statistics over (B, K, S))
    for m_g, m_r in zip(sa_g.mlp_bns, sa.mlp_bns):
        np.testing.assert_allclose(m_g.running_mean.cpu().numpy(), m_r.running_mean.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(m_g.running_var.cpu().numpy(), m_r.running_var.numpy(), rtol=1e-4, atol=1e-6)
    # one optimiser step on the GPU modules runs and changes the weights
    opt = torch.optim.Adam(list(sa_g.parameters()) + list(fp_g.parameters()), lr=1e-3)
    before = sa_g.mlp_convs[0].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, sa_g.mlp_convs[0].weight.detach())


def test_graphed_train_step_reproduces_the_eager_loop(cuda):
    """GraphedTrainStep (one CUDA graph of zero-grad + forward + loss + backward + Adam) against the eager loop of
    train_scannet_semseg.py:135-146 from the same initial weights, fixed-order backwards: identical losses and weights."""
    import torch.nn.functional as F
    from pn2_b200 import pointnet2_utils as pu
    from pn2_b200.models import GraphedTrainStep, PointNet2SemSeg
    B, N = 2, 2048
    pts = torch.from_numpy(scenes.scannet_batch(31, B, N)).to(cuda)
    xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
    rgb = pts[:, :, 3:].permute(0, 2, 1).contiguous()
    target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1

    def loss_fn(logits, tgt):
        return F.cross_entropy(logits.reshape(-1, 21), tgt.reshape(-1), ignore_index=0)

    def make():
        torch.manual_seed(3)
        net = PointNet2SemSeg(21).to(cuda).train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0  # the random stream of a replayed graph differs from the eager one
        return net, torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)

    prev = pu.set_deterministic(True)
    try:
        net_e, opt_e = make()
        eager = []
        for _ in range(3):
            opt_e.zero_grad(set_to_none=True)
            loss = loss_fn(net_e(xyz, rgb), target)
            loss.backward()
            opt_e.step()
            eager.append(float(loss.detach()))
        net_g, opt_g = make()
        stepper = GraphedTrainStep(net_g, opt_g, loss_fn, xyz, rgb, target)
        graphed = [float(stepper.step(xyz, rgb, target)) for _ in range(3)]
    finally:
        pu.set_deterministic(prev)
    assert graphed == eager, (graphed, eager)
    for (k, a), (_, b) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
        assert torch.equal(a, b), k


def test_multiview_msg_parallel_branches_match_the_sequential_chains(cuda):
    """PointNet2Multiview2Msg in training mode: geometry and image-feature chains on two streams with ONE sampling for both
    (models._MultiviewStackBase._two_branches) and the scales of every multi-scale module on streams of their own, against
    the reference's order (one chain after the other, each sampling on
    its own), eager and replayed as CUDA graphs; fixed-order backwards: identical losses and weights."""
    import torch.nn.functional as F
    from pn2_b200 import pointnet2_utils as pu
    from pn2_b200.models import GraphedTrainStep, PointNet2Multiview2Msg
    B, N = 2, 2048
    pts = torch.from_numpy(scenes.scannet_batch(41, B, N)).to(cuda)
    xyz = pts[:, :, :3].permute(0, 2, 1).contiguous()
    img = torch.randn(B, 128, N, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    target = (pts[:, :, 2].clamp(0, 2.69) / 2.7 * 20).long() + 1

    def loss_fn(logits, tgt):
        return F.cross_entropy(logits.reshape(-1, 21), tgt.reshape(-1), ignore_index=0)

    def run(parallel, graphed):
        torch.manual_seed(5)
        net = PointNet2Multiview2Msg(21).to(cuda).train()
        net.parallel_branches = parallel
        for m in net.modules():
            if hasattr(m, "parallel_scales"):
                m.parallel_scales = parallel  # the scales of every multi-scale module on streams of their own as well
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
        losses = []
        if graphed:
            stepper = GraphedTrainStep(net, opt, loss_fn, xyz, img, target)
            losses = [float(stepper.step(xyz, img, target)) for _ in range(2)]
        else:
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                loss = loss_fn(net(xyz, img), target)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
        torch.cuda.synchronize()
        return losses, {k: v.clone() for k, v in net.state_dict().items()}

    prev = pu.set_deterministic(True)
    try:
        seq = run(False, False)
        par = run(True, False)
        par_graph = run(True, True)
    finally:
        pu.set_deterministic(prev)
    for name, other in (("two streams, eager", par), ("two streams, graphed", par_graph)):
        assert other[0] == seq[0], (name, other[0], seq[0])
        for k in seq[1]:
            assert torch.equal(seq[1][k], other[1][k]), (name, k)
