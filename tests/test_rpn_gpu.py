"""Sphere IoU and NMS on the device (csrc/sphere.cu, pn2_b200/rpn.py) against the reference's own functions
(model/pointmaskrcnn.py:233-321, golden vectors in tests/golden/rpn_r1.npz) and the numpy restatement."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import rpn_ref
from pn2_b200 import rpn

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_rpn", os.path.join(HERE, "golden", "make_golden_rpn.py"))
mgr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgr)


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", list(mgr.CASES))
def test_iou_and_nms_match_reference_golden(cuda, name):
    gold = np.load(os.path.join(HERE, "golden", "rpn_r1.npz"))
    spheres, scores = mgr.rpn_inputs(name)
    s, sc = torch.from_numpy(spheres).to(cuda), torch.from_numpy(scores).to(cuda)
    iou = rpn.iou_spheres(s, s, no_grad=True).cpu().numpy()
    np.testing.assert_allclose(iou, gold[name + "/iou"], rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(iou, rpn_ref.iou_spheres(spheres, spheres))   # bit-identical to the fp32 restatement
    keep = rpn.nms(s, sc, threshold=mgr.CASES[name][2])
    assert keep.dtype == torch.int64
    np.testing.assert_array_equal(keep.cpu().numpy(), gold[name + "/keep"])


def test_batched_nms_with_counts_ties_and_empty(cuda):
    rng = np.random.default_rng(3)
    B, N = 5, 300
    spheres = np.concatenate([rng.uniform(-4, 4, (B, N, 3)), rng.uniform(0.3, 2.0, (B, N, 1))], 2).astype(np.float32)
    scores = rng.integers(0, 20, (B, N)).astype(np.float32) / 20  # many equal scores: lower index first
    counts = np.array([300, 1, 0, 150, 299], np.int32)
    keep, count = rpn.nms_batched(torch.from_numpy(spheres).to(cuda), torch.from_numpy(scores).to(cuda), 0.25,
                                  torch.from_numpy(counts).to(cuda))
    keep, count = keep.cpu().numpy(), count.cpu().numpy()
    for b in range(B):
        want = rpn_ref.nms(spheres[b, :counts[b]], scores[b, :counts[b]], 0.25)
        assert count[b] == len(want)
        np.testing.assert_array_equal(keep[b, :count[b]], want)
        assert (keep[b, count[b]:] == -1).all()
    # rectangular IoU table
    a, c = torch.from_numpy(spheres[0, :70]).to(cuda), torch.from_numpy(spheres[1, :33]).to(cuda)
    np.testing.assert_array_equal(rpn.iou_spheres(a, c).cpu().numpy(), rpn_ref.iou_spheres(spheres[0, :70], spheres[1, :33]))
