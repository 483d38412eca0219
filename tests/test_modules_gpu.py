"""GPU parity of the fused SA / FP blocks and the whole PointNet2SemSeg forward against the CPU
restatement of the reference's modules (oracle/modules_ref.py) on the same seeded inputs and weights.

Tolerance (BASELINE.json north_star): fp32 MLP outputs within 1e-5 relative.  "Relative" is taken
against the largest magnitude of the tensor (the outputs pass through ReLU / max, so many entries
are exactly 0 and an element-wise ratio is undefined): max|a-b| <= 1e-5 * max|b|, plus an
element-wise check at rtol 1e-4 / atol 1e-5*max|b|.
"""
import copy

import numpy as np
import pytest
import torch

from oracle import modules_ref
from pn2_b200 import scenes
from pn2_b200.models import PointNet2Backbone, PointNet2Multiview2, PointNet2Multiview2Msg, PointNet2SemSeg
from pn2_b200.pointnet_util import (PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg)

pytestmark = pytest.mark.gpu

REL = 1e-5     # fp32 path (BASELINE.json north_star: within 1e-5 relative)
REL_BF16 = 2e-2  # bf16 tensor-core path (north_star: within 2e-2 relative)


@pytest.fixture(params=["fp32", "bf16"])
def precision(request):
    from pn2_b200 import pointnet_util
    prev = pointnet_util.set_mlp_precision(request.param)
    yield request.param
    pointnet_util.set_mlp_precision(prev)


def tol(precision, fp32=REL):
    return fp32 if precision == "fp32" else REL_BF16


def assert_close(got, want, rel=REL):
    got, want = got.detach().cpu().numpy(), want.detach().cpu().numpy()
    assert got.shape == want.shape
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= rel * scale, "max abs err %.3e vs %.1e * %.3e" % (err, rel, scale)
    if rel <= 1e-4:
        np.testing.assert_allclose(got, want, rtol=10 * rel, atol=rel * scale)


def randomize_bn(mod, seed):
    """Non-trivial running statistics / affine so that BN folding is really exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def scene_batch(B, N, seed, D):
    pts = scenes.scannet_batch(seed, B, N)
    xyz = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1).contiguous()
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, D, N, generator=g) if D else None
    return xyz, feat


SA_CASES = [(512, 0.1, 32, 3, [32, 32, 64], 2048), (128, 0.2, 32, 64, [64, 64, 128], 512), (16, 0.8, 32, 256, [256, 256, 512], 64),
            (64, 0.4, 16, 0, [16, 32], 300), (40, 0.3, 64, 5, [24, 21], 1000), (8, 0.5, 128, 2, [8], 700),
            (32, 0.3, 8, 4, [16, 40], 500), (16, 0.5, 1, 3, [8, 8], 200), (24, 0.4, 2, 13, [32, 160], 400)]


@pytest.mark.parametrize("npoint,radius,nsample,D,mlp,N", SA_CASES)
def test_sa_fused_matches_reference_modules(cuda, precision, npoint, radius, nsample, D, mlp, N):
    torch.manual_seed(npoint + N)
    mod = PointNetSetAbstraction(npoint, radius, nsample, D + 3, mlp, False).eval()
    randomize_bn(mod, N)
    xyz, feat = scene_batch(2, N, 50 + N, D)
    with torch.no_grad():
        want_xyz, want = modules_ref.sa_forward_ref(mod, xyz, feat)
        g = copy.deepcopy(mod).to(cuda)
        got_xyz, got = g(xyz.to(cuda), feat.to(cuda) if feat is not None else None)
    np.testing.assert_array_equal(got_xyz.cpu().numpy(), want_xyz.numpy())
    assert got.is_contiguous() and got.shape == want.shape
    assert_close(got, want, tol(precision))


def test_sa_msg_fused_matches_reference_modules(cuda, precision):
    torch.manual_seed(5)
    mod = PointNetSetAbstractionMsg(128, [0.1, 0.2], [16, 32], 6, [[16, 16, 32], [32, 32, 64]]).eval()
    randomize_bn(mod, 5)
    xyz, feat = scene_batch(2, 2048, 77, 6)
    with torch.no_grad():
        want_xyz, want = modules_ref.sa_msg_forward_ref(mod, xyz, feat)
        got_xyz, got = copy.deepcopy(mod).to(cuda)(xyz.to(cuda), feat.to(cuda))
    np.testing.assert_array_equal(got_xyz.cpu().numpy(), want_xyz.numpy())
    assert_close(got, want, tol(precision))


FP_CASES = [(768, [256, 256], 64, 16, 256, 512), (131, [128, 128, 128], 4096, 512, 3, 128), (128, [128, 64], 500, 100, 0, 128),
            (20, [32], 300, 1, 4, 16),
            (2500, [128, 128], 80, 16, 452, 2048)]  # fp32 kernel: 16-row tiles, ONE 200 KB activation buffer (in-place layers)


@pytest.mark.parametrize("cin,mlp,N,S,D1,D2", FP_CASES)
def test_fp_fused_matches_reference_modules(cuda, precision, cin, mlp, N, S, D1, D2):
    torch.manual_seed(N + S)
    mod = PointNetFeaturePropagation(cin, mlp).eval()
    randomize_bn(mod, S)
    xyz1, p1 = scene_batch(2, N, 90 + N, D1)
    xyz2 = xyz1[:, :, :S].contiguous()
    p2 = torch.randn(2, D2, S)
    with torch.no_grad():
        want = modules_ref.fp_forward_ref(mod, xyz1, xyz2, p1, p2)
        got = copy.deepcopy(mod).to(cuda)(xyz1.to(cuda), xyz2.to(cuda), p1.to(cuda) if p1 is not None else None, p2.to(cuda))
    assert_close(got, want, tol(precision))


def test_unfused_training_path_matches_fused(cuda, precision):
    """The training-mode composition (our geometry kernels + torch conv/BN) in eval() with grad enabled
    must agree with the fused path, and gradients must flow to the input features."""
    torch.manual_seed(1)
    mod = PointNetSetAbstraction(64, 0.3, 32, 8 + 3, [16, 32], False).to(cuda).eval()
    xyz, feat = scene_batch(2, 1024, 13, 8)
    xyz, feat = xyz.to(cuda), feat.to(cuda)
    with torch.no_grad():
        _, fused = mod(xyz, feat)
    feat_g = feat.clone().requires_grad_(True)
    _, unfused = mod(xyz, feat_g)
    assert_close(fused, unfused, tol(precision))
    unfused.sum().backward()
    assert feat_g.grad is not None and torch.isfinite(feat_g.grad).all() and feat_g.grad.abs().sum() > 0
    fp = PointNetFeaturePropagation(8 + 32, [16]).to(cuda).train()
    feat_g.grad = None
    out = fp(xyz, xyz[:, :, :64].contiguous(), feat_g, unfused.detach())
    out.mean().backward()
    assert feat_g.grad is not None and torch.isfinite(feat_g.grad).all()


@pytest.mark.parametrize("B,N", [(2, 8192), (3, 2048)])
def test_semseg_forward_matches_reference_composition(cuda, precision, B, N):
    """Config 1 (PointNet++ SSG semseg forward, ScanNet-shaped scenes, batch 2, 8192 points) end to end."""
    torch.manual_seed(0)
    model = PointNet2SemSeg(21).eval()
    randomize_bn(model, 3)
    pts = torch.from_numpy(scenes.scannet_batch(0, B, N)).permute(0, 2, 1).contiguous()  # (B, 6, N) as the train script feeds it
    xyz, rgb = pts[:, :3, :], pts[:, 3:, :]
    want = modules_ref.semseg_forward_ref(model, xyz, rgb)
    g = copy.deepcopy(model).to(cuda)
    with torch.no_grad():
        got = g(xyz.to(cuda), rgb.to(cuda))
    assert got.shape == (B, N, 21)
    assert_close(got, want, tol(precision, 2e-5))  # 13 fused layers deep; accumulated fp32 rounding, still ~1e-5
    # default-initialised BN (running stats 0/1) as SURVEY.md 8d config 1 prescribes
    model2 = PointNet2SemSeg(21).eval()
    want2 = modules_ref.semseg_forward_ref(model2, xyz, rgb)
    with torch.no_grad():
        got2 = copy.deepcopy(model2).to(cuda)(xyz.to(cuda), rgb.to(cuda))
    assert_close(got2, want2, tol(precision, 2e-5))


def test_predict_labels_equal_argmax_of_logits(cuda, precision):
    """predict() = np.argmax over the logits (train_scannet_semseg.py:204-205).  On the tensor-core path the arg-max is
    fused into the head kernel (PN2_FLAG_OUT_ARGMAX): same logits, so the labels agree exactly, ties -> first class."""
    torch.manual_seed(1)
    model = PointNet2SemSeg(21).eval()
    randomize_bn(model, 5)
    with torch.no_grad():
        # classes 2 and 7 share weights and bias: their logits tie bit for bit, the prediction must never be 7
        model.conv2.weight[7] = model.conv2.weight[2]
        model.conv2.bias[7] = model.conv2.bias[2]
    pts = torch.from_numpy(scenes.scannet_batch(3, 2, 4096)).permute(0, 2, 1).contiguous()
    g = model.to(cuda)
    xyz, rgb = pts[:, :3, :].to(cuda), pts[:, 3:, :].to(cuda)
    with torch.no_grad():
        logits = g(xyz, rgb)
        labels = g.predict(xyz, rgb)
    assert labels.shape == (2, 4096) and labels.dtype == torch.uint8
    np.testing.assert_array_equal(labels.cpu().numpy(), np.argmax(logits.cpu().numpy(), axis=2))
    assert not (labels == 7).any()
    if precision == "bf16":
        assert g.can_fuse_labels()


@pytest.mark.parametrize("N,S,D1,D2,classes", [(3000, 300, 6, 128, 21), (1000, 64, 0, 64, 40), (777, 1, 8, 32, 200)])
def test_fused_argmax_output_of_an_fp_block(cuda, N, S, D1, D2, classes):
    """PN2_FLAG_OUT_ARGMAX on random features (predictions vary from point to point): uint8 labels equal the first
    maximum of the block's own fp32 output; the ReLU outputs tie at 0 often, which exercises the first-index rule."""
    from pn2_b200 import pointnet_util
    from pn2_b200.pointnet_util import to_channel_last
    prev = pointnet_util.set_mlp_precision("bf16")
    try:
        torch.manual_seed(N)
        mod = PointNetFeaturePropagation(D1 + D2, [64, classes]).eval()
        randomize_bn(mod, S)
        mod = mod.to(cuda)
        xyz1, p1 = scene_batch(2, N, 7 + N, D1)
        xyz2 = xyz1[:, :, :S].contiguous()
        p2 = torch.randn(2, D2, S)
        a = [to_channel_last(t.to(cuda)) if t is not None else None for t in (xyz1, xyz2, p1, p2)]
        with torch.no_grad():
            dense = mod.forward_cl(*a)
            labels = mod.forward_cl(*a, out_dtype=torch.uint8)
        assert labels.shape == (2, N) and labels.dtype == torch.uint8
        want = np.argmax(dense.cpu().numpy(), axis=2)
        np.testing.assert_array_equal(labels.cpu().numpy(), want)
        assert len(np.unique(want)) > 3
    finally:
        pointnet_util.set_mlp_precision(prev)


def test_backbone_forward_nuscenes_shape(cuda, precision):
    torch.manual_seed(2)
    model = PointNet2Backbone().eval()
    xyz_np, feat_np = scenes.lidar_sweep(0, 8192)
    xyz = torch.from_numpy(xyz_np.T.copy())[None]
    feat = torch.from_numpy(feat_np.T.copy())[None]
    want = modules_ref.backbone_forward_ref(model, xyz, feat)
    with torch.no_grad():
        got = copy.deepcopy(model).to(cuda)(xyz.to(cuda), feat.to(cuda))
    assert_close(got, want, tol(precision, 2e-5))


@pytest.mark.parametrize("cls", [PointNet2Multiview2, PointNet2Multiview2Msg])
def test_multiview_point_branch_matches_reference_composition(cuda, precision, cls):
    """BASELINE configs 2 / 3: the MSG semseg stack and the multi-view stack on lifted image features."""
    torch.manual_seed(4)
    model = cls(21).eval()
    randomize_bn(model, 9)
    B, N = 2, 8192
    xyz, img = scene_batch(B, N, 300, 128)
    want = modules_ref.multiview_stack_forward_ref(model, xyz, img)
    with torch.no_grad():
        got = copy.deepcopy(model).to(cuda)(xyz.to(cuda), img.to(cuda))
    assert got.shape == (B, N, 21)
    assert_close(got, want, tol(precision, 2e-5))


def test_multiview_forward_views_end_to_end(cuda):
    """Lifting + point branch in one call equals lifting with the oracle followed by the reference composition."""
    from oracle import oracle as orc
    from pn2_b200 import projection
    torch.manual_seed(5)
    model = PointNet2Multiview2(21).eval()
    B, N, V = 1, 8192, 3
    x, _ = scenes.scannet_scene(500, N)
    feats, depth, poses = scenes.multiview_inputs(500, x, V, 128)
    xyz = torch.from_numpy(x.T.copy())[None]
    t = lambda a: torch.from_numpy(a)[None].to(cuda)
    intr = torch.from_numpy(scenes.SCANNET_INTRINSIC)
    dmin, dmax = scenes.SCANNET_DEPTH_RANGE
    from pn2_b200 import pointnet_util
    prev = pointnet_util.set_mlp_precision("fp32")
    try:
        with torch.no_grad():
            got = copy.deepcopy(model).to(cuda).forward_views(xyz.to(cuda), t(feats), t(depth), t(poses), intr, dmin, dmax,
                                                              scenes.SCANNET_IMAGE_DIMS, scenes.SCANNET_ACCURACY)
            lifted = projection.lift_views(t(x), t(feats), t(depth), t(poses), intr, dmin, dmax, scenes.SCANNET_IMAGE_DIMS,
                                           scenes.SCANNET_ACCURACY, reduce="first")
    finally:
        pointnet_util.set_mlp_precision(prev)
    want = modules_ref.multiview_stack_forward_ref(model, xyz, lifted.cpu())
    assert_close(got, want, 2e-5)


def test_backbone_nuscenes_16384_points(cuda, precision):
    """BASELINE config 4 at the loader's real size (16 384 points): cluster FPS with 16 points per thread, 4096 centroids."""
    torch.manual_seed(6)
    model = PointNet2Backbone().eval()
    xyz_np, feat_np = scenes.lidar_sweep(1, 16384)
    xyz = torch.from_numpy(xyz_np.T.copy())[None]
    feat = torch.from_numpy(feat_np.T.copy())[None]
    want = modules_ref.backbone_forward_ref(model, xyz, feat)
    with torch.no_grad():
        got = copy.deepcopy(model).to(cuda)(xyz.to(cuda), feat.to(cuda))
    assert_close(got, want, tol(precision, 2e-5))


def test_backbone_nuscenes_raw_sweep_size_batch2(cuda, precision):
    """BASELINE config 4 at its stress size: two synthetic 32-beam sweeps resampled to 34 720 points (with replacement, so
    duplicate returns and FPS ties occur), cluster / DSMEM FPS with 4096 centroids, cell-list ball query and 3-NN."""
    torch.manual_seed(8)
    model = PointNet2Backbone().eval()
    randomize_bn(model, 11)
    n = 34720
    sweeps = [scenes.lidar_sweep(20 + i, n) for i in range(2)]
    xyz = torch.from_numpy(np.stack([s[0][:n].T for s in sweeps]).copy())
    feat = torch.from_numpy(np.stack([s[1][:n].T for s in sweeps]).copy())
    want = modules_ref.backbone_forward_ref(model, xyz, feat)
    with torch.no_grad():
        got = copy.deepcopy(model).to(cuda)(xyz.to(cuda), feat.to(cuda))
    assert got.shape == (2, 128, n)
    assert_close(got, want, tol(precision, 2e-5))


def test_bf16_logits_match_fp32_logits(cuda):
    """forward_fused(logits_dtype=torch.bfloat16): the same logits rounded once to bf16 (odd 42-byte row pitch)."""
    from pn2_b200 import pointnet_util
    prev = pointnet_util.set_mlp_precision("bf16")
    try:
        torch.manual_seed(3)
        model = PointNet2SemSeg(21).eval()
        randomize_bn(model, 7)
        model = model.to(cuda)
        pts = torch.from_numpy(scenes.scannet_batch(5, 2, 4096)).permute(0, 2, 1).contiguous().to(cuda)
        with torch.no_grad():
            full = model.forward_fused(pts[:, :3], pts[:, 3:])
            half = model.forward_fused(pts[:, :3], pts[:, 3:], logits_dtype=torch.bfloat16)
        assert half.dtype == torch.bfloat16 and half.shape == full.shape
        assert torch.equal(half, full.to(torch.bfloat16))
    finally:
        pointnet_util.set_mlp_precision(prev)
