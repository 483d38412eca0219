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
