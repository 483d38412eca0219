"""Evaluation voxelisation on the device (pn2_voxel_first_index, pn2_b200/pc_util.py) against the reference's own
utils/pc_util.py:39-51 (golden vectors) and the restated evaluation counters (train_scannet_semseg.py:225-239)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import pc_util_ref
from pn2_b200 import pc_util, scenes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_pc_util", os.path.join(HERE, "golden", "make_golden_pc_util.py"))
mgp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgp)


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", list(mgp.CASES))
def test_single_cloud_matches_reference_golden(cuda, name):
    gold = np.load(os.path.join(HERE, "golden", "pc_util_r1.npz"))
    pts, label = mgp.pc_util_inputs(name)
    uvidx, uvlabel, nvox = pc_util.point_cloud_label_to_surface_voxel_label_fast(
        torch.from_numpy(pts).to(cuda), torch.from_numpy(label).to(cuda), res=mgp.CASES[name][2])
    np.testing.assert_array_equal(uvidx.cpu().numpy(), gold[name + "/uvidx"])
    np.testing.assert_array_equal(uvlabel.cpu().numpy(), gold[name + "/uvlabel"])
    np.testing.assert_array_equal(nvox.cpu().numpy(), gold[name + "/nvox"][:3])


def test_batched_masked_counters_match_the_evaluation_loop(cuda):
    B, N, C = 4, 8192, 21
    rng = np.random.default_rng(0)
    pts = scenes.scannet_batch(900, B, N).astype(np.float32)        # (B, N, 6)
    target = rng.integers(0, C, (B, N)).astype(np.int64)
    pred = np.where(rng.random((B, N)) < 0.6, target, rng.integers(0, C, (B, N))).astype(np.int64)
    weights = (rng.random((B, N)) < 0.8).astype(np.float32) * rng.random((B, N)).astype(np.float32)
    weights[3] = 0.0                                                 # a scene with no labelled point at all
    want = pc_util_ref.voxel_accuracy_counts(pts[:3], target[:3], pred[:3], weights[:3], C, res=0.02)  # numpy would raise on scene 3
    got = pc_util.voxel_accuracy_counts(torch.from_numpy(pts).to(cuda), torch.from_numpy(target).to(cuda),
                                        torch.from_numpy(pred).to(cuda), torch.from_numpy(weights).to(cuda), C, res=0.02)
    for k, v in want.items():
        np.testing.assert_array_equal(got[k].cpu().numpy(), np.asarray(v), err_msg=k)
    assert int(want["total_seen_vox"]) > 1000
    # the fused counters (pn2_label_counts): voxel-wise as above, point-wise as train_scannet_semseg.py:210-223
    ctr = pc_util.EvalCounters(C, cuda, res=0.02)
    for _ in range(2):  # accumulates over batches
        ctr.update(torch.from_numpy(pts).to(cuda), torch.from_numpy(target).to(cuda), torch.from_numpy(pred).to(cuda).to(torch.uint8),
                   torch.from_numpy(weights).to(cuda))
    res = ctr.result()
    for k, v in want.items():
        np.testing.assert_array_equal(np.asarray(res[k]), 2 * np.asarray(v), err_msg=k)
    w = weights > 0
    assert res["total_correct"] == 2 * int(np.sum((pred == target) & (target > 0) & w))
    assert res["total_seen"] == 2 * int(np.sum((target > 0) & w))
    for l in range(C):
        assert res["total_seen_class"][l] == 2 * int(np.sum((target == l) & w))
        assert res["total_correct_class"][l] == 2 * int(np.sum((pred == l) & (target == l) & w))
        assert res["total_union_class"][l] == 2 * int(np.sum(((pred == l) | (target == l)) & w))


def test_first_index_and_edge_cases(cuda):
    # coincident points: one voxel, first index 0; a mask that drops the first points moves the first index
    p = torch.zeros(2, 16, 3, device=cuda) + 0.25
    uvidx, first, count, nvox = pc_util.voxel_first_index(p, None, 0.1)
    assert count.tolist() == [1, 1] and first[:, 0].tolist() == [0, 0] and (first[:, 1:] == -1).all()
    assert (nvox == 0).all() and (uvidx[:, 0] == 0).all()
    mask = torch.ones(2, 16, dtype=torch.bool, device=cuda)
    mask[1, :5] = False
    _, first, count, _ = pc_util.voxel_first_index(p, mask, 0.1)
    assert first[:, 0].tolist() == [0, 5]
    # per-cloud results do not depend on the other clouds of the batch
    pts = torch.from_numpy(scenes.scannet_batch(5, 3, 2048).astype(np.float32)).to(cuda)
    full = pc_util.voxel_first_index(pts, None, 0.05)
    for b in range(3):
        one = pc_util.voxel_first_index(pts[b:b + 1], None, 0.05)
        for x, y in zip(full, one):
            assert torch.equal(x[b], y[0])
    # empty batch
    e = pc_util.voxel_first_index(torch.zeros(0, 8, 3, device=cuda), None, 0.1)
    assert e[2].numel() == 0
