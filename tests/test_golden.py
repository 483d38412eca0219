"""Pins the CPU oracle to golden vectors produced by the REFERENCE's own CUDA kernels on a B200
(tests/golden/ref_cuda_r1.npz, made by tests/golden/make_golden.py).  Runs without a GPU."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
NPZ = os.path.join(HERE, "golden", "ref_cuda_r1.npz")
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

pytestmark = pytest.mark.skipif(not os.path.exists(NPZ), reason="golden fixture not generated yet")


@pytest.fixture(scope="module")
def gold():
    return np.load(NPZ)


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_matches_reference_kernels(gold, name):
    xyz, feats, M, radius = mg.golden_inputs(name)
    fps = orc.furthest_point_sample(xyz, M)
    np.testing.assert_array_equal(fps, gold[name + "/fps"])
    new_xyz = np.ascontiguousarray(orc.gather_operation(xyz.transpose(0, 2, 1), fps).transpose(0, 2, 1))
    ball = orc.ball_query(radius, 16, xyz, new_xyz)
    np.testing.assert_array_equal(ball, gold[name + "/ball"])
    grouped = orc.grouping_operation(feats, ball)
    np.testing.assert_allclose(grouped.astype(np.float64).sum((2, 3)), gold[name + "/grouped_sum"], rtol=1e-12, atol=1e-9)
    dist, nn_idx = orc.three_nn(xyz, new_xyz)
    np.testing.assert_array_equal(nn_idx, gold[name + "/nn_idx"])
    np.testing.assert_array_equal(dist, gold[name + "/nn_dist"])
    w = orc.fp_weights(dist)
    np.testing.assert_allclose(w, gold[name + "/weight"], rtol=1e-6, atol=1e-9)
    interp = orc.three_interpolate(np.ascontiguousarray(feats[:, :, :M]), nn_idx, gold[name + "/weight"])
    np.testing.assert_array_equal(interp, gold[name + "/interp"])


# ---- evaluation voxelisation: golden vectors from the reference's own utils/pc_util.py (imported, pure numpy) ----
PC_NPZ = os.path.join(HERE, "golden", "pc_util_r1.npz")
spec_pc = importlib.util.spec_from_file_location("make_golden_pc_util", os.path.join(HERE, "golden", "make_golden_pc_util.py"))
mgp = importlib.util.module_from_spec(spec_pc)
spec_pc.loader.exec_module(mgp)


@pytest.mark.parametrize("name", list(mgp.CASES))
def test_pc_util_restatement_matches_reference(name):
    from oracle import pc_util_ref
    gold = np.load(PC_NPZ)
    pts, label = mgp.pc_util_inputs(name)
    uvidx, uvlabel, nvox = pc_util_ref.point_cloud_label_to_surface_voxel_label_fast(pts, label, res=mgp.CASES[name][2])
    np.testing.assert_array_equal(uvidx, gold[name + "/uvidx"])
    np.testing.assert_array_equal(uvlabel, gold[name + "/uvlabel"])
    np.testing.assert_array_equal(nvox, gold[name + "/nvox"])
    assert uvidx.dtype == np.float32  # the reference computes the voxel index in float32


# ---- proposal-layer geometry: golden vectors from the reference's own iou_spheres / nms (torch on the CPU) ----
RPN_NPZ = os.path.join(HERE, "golden", "rpn_r1.npz")
spec_rpn = importlib.util.spec_from_file_location("make_golden_rpn", os.path.join(HERE, "golden", "make_golden_rpn.py"))
mgr = importlib.util.module_from_spec(spec_rpn)
spec_rpn.loader.exec_module(mgr)


@pytest.mark.parametrize("name", list(mgr.CASES))
def test_rpn_restatement_matches_reference(name):
    from oracle import rpn_ref
    gold = np.load(RPN_NPZ)
    spheres, scores = mgr.rpn_inputs(name)
    iou = rpn_ref.iou_spheres(spheres, spheres)
    np.testing.assert_allclose(iou, gold[name + "/iou"], rtol=2e-6, atol=1e-7)  # torch.norm's summation order is its own
    assert ((iou > 0) == (gold[name + "/iou"] > 0)).all()
    np.testing.assert_array_equal(rpn_ref.nms(spheres, scores, mgr.CASES[name][2]), gold[name + "/keep"])


# ---- multi-view lifting (rows A12-A14, N2): golden vectors from the reference's own utils/projection.py and the view
# reductions of model/pointnet2multiview.py, run on the host (tests/golden/make_golden_projection.py) ----
from lifting_cases import PROJ, boundary_cases, mgj, ref_view_params  # noqa: E402


@pytest.mark.parametrize("name", list(mgj.CASES))
@pytest.mark.parametrize("reduce", ["max", "first"])
def test_lifting_oracle_matches_reference(name, reduce):
    """The C oracle (fixed fma order), fed with the REFERENCE's per-view parameters, against the reference's decisions and
    lifted features.  Any disagreeing (view, point) must be a listed rounding-boundary case; the features must be identical
    bit for bit (sha256 of the whole map when every decision agrees)."""
    gold = np.load(PROJ)
    _, B, N, V, C = mgj.CASES[name]
    xyz, feats, depth, _ = mgj.lifting_inputs(name)
    w2c, c2, c4, nrm = ref_view_params(gold, name)
    out, pix = orc.lift_views(xyz, feats, depth, w2c, c2, c4, nrm, mgj_intr(), mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.ACCURACY, reduce)
    want = gold[name + "/pix"]
    diff = np.argwhere(pix != want)
    assert len(diff) <= 1e-3 * pix.size
    assert boundary_cases(xyz, depth, w2c, c2, c4, nrm, diff).all(), "a decision differs away from any rounding boundary"
    for b in range(B):
        np.testing.assert_array_equal(out[b][:, :mgj.KEEP_POINTS][:, (pix[b] == want[b]).all(0)[:mgj.KEEP_POINTS]],
                                      gold["%s/%s/head/%d" % (name, reduce, b)][:, (pix[b] == want[b]).all(0)[:mgj.KEEP_POINTS]])
        if (pix[b] == want[b]).all():
            assert mgj.sha(out[b]) == str(gold["%s/%s/sha/%d" % (name, reduce, b)])
    assert (want >= 0).mean() > 0.15


def mgj_intr():
    return np.array([mgj.INTRINSIC[0][0], mgj.INTRINSIC[1][1], mgj.INTRINSIC[0][2], mgj.INTRINSIC[1][2]], np.float32)


@pytest.mark.parametrize("name", list(mgj.CASES))
def test_projection_restatement_matches_reference(name):
    """oracle/projection_ref.py (the torch-op restatement the GPU agreement test uses) against the reference's vectors:
    view parameters to the last bit on the same BLAS, decisions identical up to listed boundary cases, both reductions."""
    import torch
    from oracle import projection_ref
    gold = np.load(PROJ)
    _, B, N, V, C = mgj.CASES[name]
    xyz, feats, depth, poses = mgj.lifting_inputs(name)
    w2c, c2, c4, nrm = ref_view_params(gold, name)
    for b in range(B):
        pix = []
        for v in range(V):
            c2w = torch.from_numpy(poses[b, v])
            cc = projection_ref.corners_ref(mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS, c2w)
            np.testing.assert_allclose(cc[:, :, 0].numpy(), gold[name + "/corners"][b, v], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(projection_ref.normals_ref(cc).numpy(), gold[name + "/normals"][b, v], rtol=1e-5, atol=1e-6)
            pix.append(projection_ref.compute_projection_ref(torch.from_numpy(xyz[b]), torch.from_numpy(depth[b, v]), c2w,
                                                             mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS, mgj.ACCURACY))
        got = np.stack([p.numpy() for p in pix])
        want = gold[name + "/pix"][b]
        diff = np.argwhere(got != want)
        diff = np.concatenate([np.full((len(diff), 1), b), diff], 1)
        assert len(diff) <= 1e-3 * got.size and boundary_cases(xyz, depth, w2c, c2, c4, nrm, diff).all()
        if len(diff) == 0:
            for reduce in ("max", "first"):
                o = projection_ref.lift_ref(torch.from_numpy(feats[b]), pix, reduce).numpy()
                assert mgj.sha(o) == str(gold["%s/%s/sha/%d" % (name, reduce, b)])


def test_frustum_count_oracle_matches_reference():
    """Row N2: the loader's best-view selection (data_utils/ScanNetDataLoader.py:87-105).  The fp64 oracle, fed with the
    reference's fp32 corners / normals, must give the reference's membership masks exactly (fp64 leaves no boundary cases)."""
    gold = np.load(PROJ)
    x, poses = mgj.frustum_inputs()
    counts, mask = orc.frustum_count(x, gold["frustum/corners"], gold["frustum/normals"], return_mask=True)
    np.testing.assert_array_equal(counts, gold["frustum/counts"])
    np.testing.assert_array_equal(np.packbits(mask, axis=-1), gold["frustum/masks"])
    assert (counts == 0).sum() >= 5 and counts.max() > 5000


# ---- SA / SA-MSG / FP modules and the networks built from them (rows A8-A11): golden vectors from the reference's own,
# unmodified classes run on the host over the oracle-backed pointnet2_cuda (tests/golden/make_golden_modules.py) ----
from module_cases import MODS, check_sample, mgm, unit_seed  # noqa: E402


def _t(a):
    import torch
    return None if a is None else torch.from_numpy(a)


@pytest.mark.parametrize("name", sorted(mgm.UNIT_CASES))
def test_module_restatement_matches_reference_classes(name):
    """oracle/modules_ref.py (what the GPU tests and the CPU baseline use) against the reference's own classes."""
    import torch
    from oracle import modules_ref
    from oracle.seeded import fill_seeded
    from pn2_b200 import pointnet_util as pu
    gold = np.load(MODS)
    kind, args, B, N, D, extra = mgm.UNIT_CASES[name]
    ctor = {"sa": pu.PointNetSetAbstraction, "msg": pu.PointNetSetAbstractionMsg, "fp": pu.PointNetFeaturePropagation}[kind]
    mod = fill_seeded(ctor(*args), unit_seed(name)).eval()   # same constructor arguments and state_dict keys as the reference
    inp = [_t(a) for a in mgm.unit_inputs(name)]
    with torch.no_grad():
        if kind == "fp":
            out = modules_ref.fp_forward_ref(mod, *inp)
        else:
            new_xyz, out = (modules_ref.sa_forward_ref if kind == "sa" else modules_ref.sa_msg_forward_ref)(mod, *inp)
            np.testing.assert_array_equal(new_xyz.numpy(), gold[name + "/new_xyz"])
        want = gold[name + "/out"]
        np.testing.assert_allclose(out.numpy(), want, rtol=1e-5, atol=2e-6 * np.abs(want).max())
        if name == "sa_ssg":
            mod.train()
            _, out = modules_ref.sa_forward_ref(mod, *inp)
            want = gold[name + "/train_out"]
            np.testing.assert_allclose(out.numpy(), want, rtol=1e-4, atol=5e-6 * np.abs(want).max())
            np.testing.assert_allclose(mod.mlp_bns[0].running_mean.numpy(), gold[name + "/train_running_mean0"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(mod.mlp_bns[0].running_var.numpy(), gold[name + "/train_running_var0"], rtol=1e-5, atol=1e-7)


def test_network_restatements_match_reference_classes():
    """PointNet2SemSeg at config 1 (B=2, N=8192), the nuScenes backbone at 16384 points, and the two multi-view stacks
    including the lifting, against the reference's own classes."""
    import torch
    from oracle import modules_ref, projection_ref
    from oracle.seeded import fill_seeded
    from pn2_b200 import models
    gold = np.load(MODS)
    xyz, rgb = mgm.semseg_inputs()
    net = fill_seeded(models.PointNet2SemSeg(mgm.NUM_CLASSES), 200).eval()
    y = modules_ref.semseg_forward_ref(net, _t(xyz), _t(rgb)).numpy()
    assert check_sample(y, gold, "semseg_eval", 1e-5) < 1e-5
    assert (y.argmax(-1) == gold["semseg_eval/argmax"]).mean() > 0.9999
    bx, bf = mgm.backbone_inputs(16384)
    bb = fill_seeded(models.PointNet2Backbone(), 210).eval()
    y = modules_ref.backbone_forward_ref(bb, _t(bx), _t(bf)).numpy()
    check_sample(y, gold, "backbone16k", 1e-5, axis=2, stride=16)
    for cls, tag, B, V, seed, reduce in ((models.PointNet2Multiview2, "mv2_first", 2, 3, 220, "first"),
                                         (models.PointNet2Multiview2Msg, "mv2msg_max", 1, 5, 230, "max")):
        mx, mf, md, mp = mgm.multiview_inputs(B, V)
        imgs = []
        for b in range(B):
            pts = _t(np.ascontiguousarray(mx[b].T))
            pix = [projection_ref.compute_projection_ref(pts, _t(md[b, v]), _t(mp[b, v]), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX,
                                                         mgj.IMAGE_DIMS, mgj.ACCURACY) for v in range(V)]
            img = projection_ref.lift_ref(_t(mf[b]), pix, reduce)
            assert mgj.sha(img.numpy()) == str(gold[tag + "/image_features_sha"][b])
            imgs.append(img)
        net = fill_seeded(cls(mgm.NUM_CLASSES), seed).eval()
        y = modules_ref.multiview_stack_forward_ref(net, _t(mx), torch.stack(imgs)).numpy()
        check_sample(y, gold, tag, 1e-5)
