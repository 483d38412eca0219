"""torch.library layer on the GPU: torch.ops.pn2.* return exactly what the autograd.Function surface returns (the same
kernels), their registered backwards match it too, torch.library.opcheck accepts the registrations, and a composed SA
slice traces with fullgraph=True (backend "eager": the kernels stay ours, nothing is code-generated)."""
import numpy as np
import pytest
import torch

import pn2_b200.ops  # noqa: F401  (registers torch.ops.pn2)
from pn2_b200 import pointnet2_utils as pu
from pn2_b200 import scenes

pytestmark = pytest.mark.gpu


def _inputs(cuda, B=2, N=2048, M=256, C=16, K=16):
    pts = scenes.scannet_batch(7, B, N)
    xyz = torch.from_numpy(np.ascontiguousarray(pts[:, :, :3])).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(5)
    feat = torch.randn(B, C, N, generator=g).to(cuda)
    return xyz, feat, M, K


def test_forward_identical_to_function_surface(cuda):
    xyz, feat, M, K = _inputs(cuda)
    idx = torch.ops.pn2.furthest_point_sample(xyz, M)
    assert torch.equal(idx, pu.furthest_point_sample(xyz, M))
    xyz_cf = xyz.transpose(1, 2).contiguous()
    new_xyz = torch.ops.pn2.gather_operation(xyz_cf, idx)
    assert torch.equal(new_xyz, pu.gather_operation(xyz_cf, idx))
    new_xyz = new_xyz.transpose(1, 2).contiguous()
    bq = torch.ops.pn2.ball_query(0.2, K, xyz, new_xyz)
    assert torch.equal(bq, pu.ball_query(0.2, K, xyz, new_xyz))
    assert torch.equal(torch.ops.pn2.grouping_operation(feat, bq), pu.grouping_operation(feat, bq))
    dist, i3 = torch.ops.pn2.three_nn(xyz, new_xyz)
    dist_f, i3_f = pu.three_nn(xyz, new_xyz)
    assert torch.equal(i3, i3_f) and torch.equal(dist, dist_f)
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    coarse = pu.gather_operation(feat, idx)
    assert torch.equal(torch.ops.pn2.three_interpolate(coarse, i3, w), pu.three_interpolate(coarse, i3, w))


def test_backward_identical_to_function_surface(cuda):
    xyz, feat, M, K = _inputs(cuda)
    prev = pu.set_deterministic(True)  # fixed summation order: the two surfaces must agree bit for bit
    try:
        idx = pu.furthest_point_sample(xyz, M)
        new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        bq = pu.ball_query(0.2, K, xyz, new_xyz)
        dist, i3 = pu.three_nn(xyz, new_xyz)
        w = 1.0 / (dist + 1e-8)
        w = (w / w.sum(dim=2, keepdim=True)).contiguous()

        def run(gather, group, interp):
            f = feat.clone().requires_grad_(True)
            coarse = gather(f, idx)
            y = group(f, bq).sum(dim=3) * 0.5 + coarse
            z = interp(y, i3, w)
            (z * z).sum().backward()
            return f.grad

        g_ops = run(torch.ops.pn2.gather_operation, torch.ops.pn2.grouping_operation, torch.ops.pn2.three_interpolate)
        g_fun = run(pu.gather_operation, pu.grouping_operation, pu.three_interpolate)
        assert torch.equal(g_ops, g_fun)
    finally:
        pu.set_deterministic(prev)


def test_opcheck(cuda):
    xyz, feat, M, K = _inputs(cuda, N=512, M=64, C=4, K=8)
    idx = pu.furthest_point_sample(xyz, M)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    bq = pu.ball_query(0.2, K, xyz, new_xyz)
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.pn2.furthest_point_sample.default, (xyz, M), test_utils=tests)
    torch.library.opcheck(torch.ops.pn2.ball_query.default, (0.2, K, xyz, new_xyz), test_utils=tests)
    torch.library.opcheck(torch.ops.pn2.grouping_operation.default, (feat.clone().requires_grad_(True), bq), test_utils=tests)
    torch.library.opcheck(torch.ops.pn2.gather_operation.default, (feat.clone().requires_grad_(True), idx), test_utils=tests)


def test_traces_without_graph_breaks(cuda):
    xyz, feat, M, K = _inputs(cuda, N=1024, M=128, C=8, K=16)

    def sa_slice(xyz, feat):
        idx = torch.ops.pn2.furthest_point_sample(xyz, M)
        new_xyz = torch.ops.pn2.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        bq = torch.ops.pn2.ball_query(0.2, K, xyz, new_xyz)
        return torch.ops.pn2.grouping_operation(feat, bq).max(dim=3).values

    want = sa_slice(xyz, feat)
    traced = torch.compile(sa_slice, backend="eager", fullgraph=True)  # fullgraph: any graph break raises
    assert torch.equal(traced(xyz, feat), want)
