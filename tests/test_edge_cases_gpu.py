"""Edge cases through the operator layer / C ABI: empty and ragged shapes, maximum nsample, M > N, degenerate clouds."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from pn2_b200 import Pn2Error
from pn2_b200 import pointnet2_utils as pu

pytestmark = pytest.mark.gpu


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def test_empty_batches_and_zero_queries(cuda):
    xyz = torch.zeros((0, 128, 3), device=cuda)
    assert pu.furthest_point_sample(xyz, 16).shape == (0, 16)
    x = torch.rand((2, 64, 3), device=cuda)
    q = torch.zeros((2, 0, 3), device=cuda)
    assert pu.ball_query(0.1, 8, x, q).shape == (2, 0, 8)
    d, i = pu.three_nn(q, x)
    assert d.shape == (2, 0, 3) and i.shape == (2, 0, 3)
    f = torch.rand((2, 5, 64), device=cuda)
    assert pu.gather_operation(f, torch.zeros((2, 0), dtype=torch.int32, device=cuda)).shape == (2, 5, 0)
    assert pu.grouping_operation(f, torch.zeros((2, 0, 4), dtype=torch.int32, device=cuda)).shape == (2, 5, 0, 4)


def test_fps_more_samples_than_points_and_identical_points(cuda):
    # M > N: once every running minimum is 0 the total order returns index 0
    xyz = np.random.default_rng(1).random((2, 40, 3)).astype(np.float32)
    got = pu.furthest_point_sample(dev(xyz, cuda), 100).cpu().numpy()
    np.testing.assert_array_equal(got, orc.furthest_point_sample(xyz, 100))
    assert (got[:, 40:] == 0).all()
    # a cloud of one repeated point: every distance ties at 0
    same = np.ones((1, 2048, 3), np.float32)
    np.testing.assert_array_equal(pu.furthest_point_sample(dev(same, cuda), 64).cpu().numpy(), orc.furthest_point_sample(same, 64))
    np.testing.assert_array_equal(pu.ball_query(0.5, 16, dev(same, cuda), dev(same[:, :4], cuda)).cpu().numpy(),
                                  orc.ball_query(0.5, 16, same, same[:, :4]))


def test_ball_query_large_nsample_and_zero_radius(cuda):
    xyz = np.random.default_rng(2).random((1, 3000, 3)).astype(np.float32)
    q = xyz[:, :50].copy()
    for r, K in [(0.2, 200), (0.0, 4), (1e-9, 4)]:
        np.testing.assert_array_equal(pu.ball_query(r, K, dev(xyz, cuda), dev(q, cuda)).cpu().numpy(), orc.ball_query(r, K, xyz, q))


def test_three_nn_one_and_two_known_points(cuda):
    unk = np.random.default_rng(3).random((1, 100, 3)).astype(np.float32)
    for m in (1, 2):
        known = np.random.default_rng(4).random((1, m, 3)).astype(np.float32)
        d, i = pu.three_nn(dev(unk, cuda), dev(known, cuda))
        wd, wi = orc.three_nn(unk, known)
        np.testing.assert_array_equal(i.cpu().numpy(), wi)
        np.testing.assert_array_equal(d.cpu().numpy(), wd)
        assert np.isinf(wd[..., 2]).all()


def test_contiguity_and_dtype_errors(cuda):
    x = torch.rand((2, 64, 3), device=cuda)
    with pytest.raises(Pn2Error):
        pu.furthest_point_sample(x.transpose(1, 2), 8)          # not contiguous (the reference asserts the same)
    with pytest.raises(Pn2Error):
        pu.gather_operation(torch.rand((2, 4, 64), device=cuda), torch.zeros((2, 8), dtype=torch.int64, device=cuda))
    from pn2_b200 import _lib
    lib = _lib.load()
    assert lib.pn2_sa_mlp_max(1, 8, 1, 3, 0, None, None, None, None, 0, None, None, 0, 0, None) != 0  # nsample not a power of two / null mlp
    assert len(lib.pn2_last_error()) > 0


@pytest.mark.parametrize("shape", [(2, 3, 8192), (3, 6, 1000), (1, 8, 256), (2, 8192, 3), (2, 300, 5), (2, 3, 100), (2, 64, 1024),
                                   (1, 131, 777), (2, 9, 4096)])
def test_transpose_matches_torch(cuda, shape):
    """pn2_transpose ((B, C, N) <-> (B, N, C) at the module boundary): the narrow-axis kernels (<= 8 rows or columns,
    long axis >= 256) and the tiled kernel give exactly torch's permuted copy."""
    from pn2_b200.pointnet_util import to_channel_first, to_channel_last
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to(cuda)
    assert torch.equal(to_channel_last(x), x.permute(0, 2, 1).contiguous())
    assert torch.equal(to_channel_first(x), x.permute(0, 2, 1).contiguous())
