"""GPU parity of the SA / SA-MSG / FP modules and the networks against vectors produced by the REFERENCE's own, unmodified
module classes (model/pointnet_util.py:70-221, model/pointnet2.py:131-162, model/pointmaskrcnn.py:8-32,
model/pointnet2multiview.py:61-121 / 179-233 + utils/projection.py), tests/golden/modules_r2.npz.

Tolerances (BASELINE.json north_star): fp32 MLP path 1e-5 relative (2e-5 for the 13-layer networks), bf16 tensor-core path
2e-2 relative, both against the largest magnitude of the tensor; sampled centroids (new_xyz) bit-exact."""
import numpy as np
import pytest
import torch

from lifting_cases import mgj
from module_cases import MODS, check_sample, mgm, unit_seed
from oracle.seeded import fill_seeded
from pn2_b200 import models, pointnet_util as pu, projection

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fp32", "bf16"])
def precision(request):
    prev = pu.set_mlp_precision(request.param)
    yield request.param
    pu.set_mlp_precision(prev)


def rel(precision, fp32=1e-5):
    return fp32 if precision == "fp32" else 2e-2


def _t(a, dev):
    return None if a is None else torch.from_numpy(a).to(dev)


@pytest.mark.parametrize("name", sorted(mgm.UNIT_CASES))
def test_modules_reproduce_reference_classes(cuda, precision, name):
    gold = np.load(MODS)
    kind, args, B, N, D, extra = mgm.UNIT_CASES[name]
    ctor = {"sa": pu.PointNetSetAbstraction, "msg": pu.PointNetSetAbstractionMsg, "fp": pu.PointNetFeaturePropagation}[kind]
    mod = fill_seeded(ctor(*args), unit_seed(name)).eval().to(cuda)
    inp = [_t(a, cuda) for a in mgm.unit_inputs(name)]
    with torch.no_grad():
        res = mod(*inp)
    if kind != "fp":
        np.testing.assert_array_equal(res[0].cpu().numpy(), gold[name + "/new_xyz"])
        res = res[1]
    want = gold[name + "/out"]
    got = res.cpu().numpy()
    assert got.shape == want.shape
    r = rel(precision) if name != "sa_group_all" else 1e-3   # group_all runs torch conv (cuDNN, TF32 allowed by default)
    assert np.abs(got - want).max() <= r * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()


def test_sa_train_mode_batch_statistics_reproduce_reference(cuda):
    """Training-mode BatchNorm2d (batch statistics over (B, K, S), model/pointnet_util.py:105-107) on our geometry kernels."""
    gold = np.load(MODS)
    name = "sa_ssg"
    kind, args, B, N, D, extra = mgm.UNIT_CASES[name]
    mod = fill_seeded(pu.PointNetSetAbstraction(*args), unit_seed(name)).train().to(cuda)
    inp = [_t(a, cuda) for a in mgm.unit_inputs(name)]
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            new_xyz, out = mod(*inp)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    want = gold[name + "/train_out"]
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), gold[name + "/new_xyz"])
    assert np.abs(out.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
    np.testing.assert_allclose(mod.mlp_bns[0].running_mean.cpu().numpy(), gold[name + "/train_running_mean0"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(mod.mlp_bns[0].running_var.cpu().numpy(), gold[name + "/train_running_var0"], rtol=1e-4, atol=1e-6)


def test_semseg_config1_reproduces_reference_class(cuda, precision):
    """BASELINE config 1: PointNet2SemSeg forward, B=2, N=8192, against the reference's own PointNet2SemSeg."""
    gold = np.load(MODS)
    xyz, rgb = mgm.semseg_inputs()
    net = fill_seeded(models.PointNet2SemSeg(mgm.NUM_CLASSES), 200).eval().to(cuda)
    with torch.no_grad():
        y = net(_t(xyz, cuda), _t(rgb, cuda))
        labels = net.predict(_t(xyz, cuda), _t(rgb, cuda)).cpu().numpy()
    y = y.cpu().numpy()
    check_sample(y, gold, "semseg_eval", rel(precision, 2e-5))
    agree = (labels == gold["semseg_eval/argmax"]).mean()
    assert agree >= (0.9995 if precision == "fp32" else 0.97), agree


def test_semseg_train_mode_reproduces_reference_class(cuda):
    """The training-mode forward (batch-statistics BatchNorm everywhere, dropout off) of the whole network."""
    gold = np.load(MODS)
    xyz, rgb = mgm.semseg_inputs()
    net = fill_seeded(models.PointNet2SemSeg(mgm.NUM_CLASSES), 200).to(cuda).train()
    net.drop1.eval()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y = net(_t(xyz, cuda), _t(rgb, cuda)).cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    check_sample(y, gold, "semseg_train", 1e-4)
    np.testing.assert_allclose(net.sa1.mlp_bns[0].running_mean.cpu().numpy(), gold["semseg_train/sa1_bn0_running_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(net.fp1.mlp_bns[2].running_var.cpu().numpy(), gold["semseg_train/fp1_bn2_running_var"], rtol=1e-3, atol=1e-6)


def test_backbone_config4_reproduces_reference_class(cuda, precision):
    gold = np.load(MODS)
    bx, bf = mgm.backbone_inputs(16384)
    bb = fill_seeded(models.PointNet2Backbone(), 210).eval().to(cuda)
    with torch.no_grad():
        y = bb(_t(bx, cuda), _t(bf, cuda)).cpu().numpy()
    check_sample(y, gold, "backbone16k", rel(precision, 2e-5), axis=2, stride=16)


@pytest.mark.parametrize("tag", ["mv2_first", "mv2msg_max"])
def test_multiview_networks_reproduce_reference_classes(cuda, precision, tag):
    """BASELINE configs 3 / 2: lifting (compute_projection + Projection + view reduction) and the point branch in one call,
    against the reference's PointNet2Multiview2 / PointNet2Multiview2Msg driven as its training loop drives them."""
    gold = np.load(MODS)
    cls, B, V, seed, reduce = {"mv2_first": (models.PointNet2Multiview2, 2, 3, 220, "first"),
                               "mv2msg_max": (models.PointNet2Multiview2Msg, 1, 5, 230, "max")}[tag]
    mx, mf, md, mp = mgm.multiview_inputs(B, V)
    net = fill_seeded(cls(mgm.NUM_CLASSES), seed).eval().to(cuda)
    assert net.reduce == reduce
    pts = _t(np.ascontiguousarray(mx.transpose(0, 2, 1)), cuda)
    with torch.no_grad():
        img = projection.lift_views(pts, _t(mf, cuda), _t(md, cuda), _t(mp, cuda), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX,
                                    mgj.IMAGE_DIMS, mgj.ACCURACY, reduce=reduce)
        y = net.forward_views(_t(mx, cuda), _t(mf, cuda), _t(md, cuda), _t(mp, cuda), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX,
                              mgj.IMAGE_DIMS, mgj.ACCURACY).cpu().numpy()
    img = img.cpu().numpy()
    exact = [mgj.sha(img[b]) == str(gold[tag + "/image_features_sha"][b]) for b in range(B)]
    np.testing.assert_allclose(img.astype(np.float64).sum((1, 2)), gold[tag + "/image_features_sum"], atol=60.0)
    # a point on a rounding boundary may be lifted from a neighbouring pixel (the product's view parameters differ from the
    # reference's BLAS-ordered ones in the last place); its logits then differ legitimately, so the bound is on the sample
    # as a whole when every lifted column is identical, and on all but a handful of points otherwise
    if all(exact):
        check_sample(y, gold, tag, rel(precision, 2e-5))
    else:
        want = gold[tag + "/sample"]
        err = np.abs(y[:, ::mgm.STRIDE] - want).max(-1)
        assert (err > rel(precision, 2e-5) * float(gold[tag + "/absmax"])).mean() < 0.01
