"""torch.library layer (pn2_b200/ops.py) without a GPU: every operator is registered, its fake kernel gives the
reference's output shapes and dtypes (model/pointnet2_utils.py:10-226), and CPU tensors are refused by the dispatcher
(no CPU kernel is registered -- the path has no fallback)."""
import pytest
import torch

import pn2_b200.ops as ops


def test_all_ops_registered():
    for name in ops.OPS:
        assert hasattr(torch.ops.pn2, name), name


def test_fake_kernels_give_reference_shapes():
    B, N, M, K, C = 2, 100, 16, 8, 5
    xyz = torch.empty(B, N, 3, device="meta")
    new_xyz = torch.empty(B, M, 3, device="meta")
    feat = torch.empty(B, C, N, device="meta")
    idx1 = torch.empty(B, M, dtype=torch.int32, device="meta")
    idx2 = torch.empty(B, M, K, dtype=torch.int32, device="meta")
    idx3 = torch.empty(B, N, 3, dtype=torch.int32, device="meta")
    w3 = torch.empty(B, N, 3, device="meta")
    coarse = torch.empty(B, C, M, device="meta")
    out = torch.ops.pn2.furthest_point_sample(xyz, M)
    assert out.shape == (B, M) and out.dtype == torch.int32
    assert torch.ops.pn2.gather_operation(feat, idx1).shape == (B, C, M)
    assert torch.ops.pn2.gather_operation_grad(torch.empty(B, C, M, device="meta"), idx1, N).shape == (B, C, N)
    dist, idx = torch.ops.pn2.three_nn(xyz, new_xyz)
    assert dist.shape == (B, N, 3) and idx.shape == (B, N, 3) and idx.dtype == torch.int32 and dist.dtype == torch.float32
    assert torch.ops.pn2.three_interpolate(coarse, idx3, w3).shape == (B, C, N)
    assert torch.ops.pn2.three_interpolate_grad(torch.empty(B, C, N, device="meta"), idx3, w3, M).shape == (B, C, M)
    assert torch.ops.pn2.grouping_operation(feat, idx2).shape == (B, C, M, K)
    assert torch.ops.pn2.grouping_operation_grad(torch.empty(B, C, M, K, device="meta"), idx2, N).shape == (B, C, N)
    bq = torch.ops.pn2.ball_query(0.1, K, xyz, new_xyz)
    assert bq.shape == (B, M, K) and bq.dtype == torch.int32


def test_fake_tensor_mode_traces_a_set_abstraction_slice():
    """FPS -> gather -> ball query -> grouping composes under FakeTensorMode (what torch.export / torch.compile run)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        xyz = torch.empty(2, 256, 3, device="cuda")
        feat = torch.empty(2, 6, 256, device="cuda")
        idx = torch.ops.pn2.furthest_point_sample(xyz, 32)
        new_xyz = torch.ops.pn2.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        bq = torch.ops.pn2.ball_query(0.2, 16, xyz, new_xyz)
        grouped = torch.ops.pn2.grouping_operation(feat, bq)
        assert grouped.shape == (2, 6, 32, 16) and grouped.device.type == "cuda"


def test_cpu_tensors_are_refused():
    with pytest.raises(NotImplementedError):
        torch.ops.pn2.furthest_point_sample(torch.zeros(1, 64, 3), 8)
    with pytest.raises(NotImplementedError):
        torch.ops.pn2.grouping_operation(torch.zeros(1, 3, 64), torch.zeros(1, 8, 4, dtype=torch.int32))
