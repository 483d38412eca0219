"""The CUDA kernels against the committed golden vectors of the reference's own kernels (no live reference needed)."""
import os

import numpy as np
import pytest
import torch

from pn2_b200 import pointnet2_utils as pu
from test_golden import NPZ, mg

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(NPZ), reason="golden fixture not generated yet")]


@pytest.mark.parametrize("name", list(mg.CASES))
def test_kernels_match_reference_golden(cuda, name):
    gold = np.load(NPZ)
    xyz, feats, M, radius = mg.golden_inputs(name)
    x, f = torch.from_numpy(xyz).to(cuda), torch.from_numpy(feats).to(cuda)
    fps = pu.furthest_point_sample(x, M)
    np.testing.assert_array_equal(fps.cpu().numpy(), gold[name + "/fps"])
    new_xyz = pu.gather_operation(x.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    ball = pu.ball_query(radius, 16, x, new_xyz)
    np.testing.assert_array_equal(ball.cpu().numpy(), gold[name + "/ball"])
    dist, nn_idx = pu.three_nn(x, new_xyz)
    np.testing.assert_array_equal(nn_idx.cpu().numpy(), gold[name + "/nn_idx"])
    np.testing.assert_array_equal(dist.cpu().numpy(), gold[name + "/nn_dist"])
    interp = pu.three_interpolate(f[:, :, :M].contiguous(), nn_idx, torch.from_numpy(gold[name + "/weight"]).to(cuda))
    np.testing.assert_array_equal(interp.cpu().numpy(), gold[name + "/interp"])
