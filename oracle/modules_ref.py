"""CPU restatement of the reference's SA / FP modules and semseg forward (TEST INFRASTRUCTURE ONLY).

Follows model/pointnet_util.py:20-47 (sample_and_group), :85-111 (PointNetSetAbstraction.forward),
:133-171 (PointNetSetAbstractionMsg.forward), :185-221 (PointNetFeaturePropagation.forward) and
model/pointnet2.py:147-162 (PointNet2SemSeg.forward) step by step, with the geometry operators
taken from the C oracle (oracle/pn2_oracle.c) and the 1x1 conv / BatchNorm / ReLU / max from
torch on the CPU -- i.e. "the reference's path with the device swapped for the host".  The reference
itself has no CPU path (model/pointnet2_utils.py:7 hard-imports the CUDA extension), so this is
what bench.py times as the CPU baseline (kind "port").

The functions take any nn.Module that carries the reference's attribute names (npoint, radius,
nsample, mlp_convs, mlp_bns, conv_blocks, bn_blocks, ...), living on the CPU.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def sample_and_group_ref(npoint, radius, nsample, xyz, points):
    """xyz (B,N,3), points (B,N,D) torch CPU -> new_xyz (B,S,3), new_points (B,S,K,3+D); pointnet_util.py:20-47"""
    xyz_np = _np(xyz)
    fps_idx = orc.furthest_point_sample(xyz_np, npoint)
    new_xyz = orc.gather_operation(xyz_np.transpose(0, 2, 1), fps_idx).transpose(0, 2, 1)
    new_xyz = np.ascontiguousarray(new_xyz)
    idx = orc.ball_query(radius, nsample, xyz_np, new_xyz)
    grouped_xyz = orc.grouping_operation(xyz_np.transpose(0, 2, 1), idx).transpose(0, 2, 3, 1)
    grouped_xyz_norm = grouped_xyz - new_xyz[:, :, None, :]
    if points is not None:
        grouped_points = orc.grouping_operation(_np(points).transpose(0, 2, 1), idx).transpose(0, 2, 3, 1)
        new_points = np.concatenate([grouped_xyz_norm, grouped_points], axis=-1)
    else:
        new_points = grouped_xyz_norm
    return _t(new_xyz), _t(new_points), fps_idx, idx


def sa_forward_ref(mod, xyz, points):
    """PointNetSetAbstraction.forward (pointnet_util.py:85-111): xyz (B,3,N), points (B,D,N) or None."""
    xyz_t = xyz.permute(0, 2, 1)
    pts_t = points.permute(0, 2, 1) if points is not None else None
    if mod.group_all:
        B, N, C = xyz_t.shape
        new_xyz = torch.zeros(B, 1, C)
        new_points = xyz_t.reshape(B, 1, N, C)
        if pts_t is not None:
            new_points = torch.cat([new_points, pts_t.reshape(B, 1, N, -1)], dim=-1)
    else:
        new_xyz, new_points, _, _ = sample_and_group_ref(mod.npoint, mod.radius, mod.nsample, xyz_t, pts_t)
    new_points = new_points.permute(0, 3, 2, 1)
    for conv, bn in zip(mod.mlp_convs, mod.mlp_bns):
        new_points = F.relu(bn(conv(new_points)))
    new_points = torch.max(new_points, 2)[0]
    return new_xyz.permute(0, 2, 1), new_points


def sa_msg_forward_ref(mod, xyz, points):
    """PointNetSetAbstractionMsg.forward (pointnet_util.py:133-171)."""
    xyz_np = _np(xyz.permute(0, 2, 1))
    B, N, C = xyz_np.shape
    S = mod.npoint
    fps_idx = orc.furthest_point_sample(xyz_np, S)
    new_xyz = np.ascontiguousarray(orc.gather_operation(xyz_np.transpose(0, 2, 1), fps_idx).transpose(0, 2, 1))
    outs = []
    for i, radius in enumerate(mod.radius_list):
        K = mod.nsample_list[i]
        idx = orc.ball_query(radius, K, xyz_np, new_xyz)
        grouped_xyz = orc.grouping_operation(xyz_np.transpose(0, 2, 1), idx).transpose(0, 2, 3, 1)
        grouped_xyz = grouped_xyz - new_xyz[:, :, None, :]
        if points is not None:
            gp = orc.grouping_operation(_np(points), idx).transpose(0, 2, 3, 1)
            grouped = np.concatenate([gp, grouped_xyz], axis=-1)  # features first, xyz last (:157)
        else:
            grouped = grouped_xyz
        g = _t(grouped).permute(0, 3, 2, 1)
        for conv, bn in zip(mod.conv_blocks[i], mod.bn_blocks[i]):
            g = F.relu(bn(conv(g)))
        outs.append(torch.max(g, 2)[0])
    return _t(new_xyz).permute(0, 2, 1), torch.cat(outs, dim=1)


def fp_forward_ref(mod, xyz1, xyz2, points1, points2):
    """PointNetFeaturePropagation.forward (pointnet_util.py:185-221)."""
    xyz1_np, xyz2_np = _np(xyz1.permute(0, 2, 1)), _np(xyz2.permute(0, 2, 1))
    B, N, _ = xyz1_np.shape
    S = xyz2_np.shape[1]
    if S == 1:
        interpolated = points2.permute(0, 2, 1).repeat(1, N, 1)
    else:
        dist, idx = orc.three_nn(xyz1_np, xyz2_np)
        dist = _t(dist)
        dist[dist < 1e-10] = 1e-10
        weight = 1.0 / dist
        weight = weight / torch.sum(weight, dim=-1).view(B, N, 1)
        interp = orc.three_interpolate(_np(points2), idx, _np(weight))
        interpolated = _t(interp).permute(0, 2, 1)
    if points1 is not None:
        new_points = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)
    else:
        new_points = interpolated
    new_points = new_points.permute(0, 2, 1)
    for conv, bn in zip(mod.mlp_convs, mod.mlp_bns):
        new_points = F.relu(bn(conv(new_points)))
    return new_points


def semseg_forward_ref(model, xyz, points):
    """PointNet2SemSeg.forward (model/pointnet2.py:147-162) on the CPU: xyz (B,3,N), points (B,D,N) -> (B,N,classes)"""
    with torch.no_grad():
        l1_xyz, l1_points = sa_forward_ref(model.sa1, xyz, points)
        l2_xyz, l2_points = sa_forward_ref(model.sa2, l1_xyz, l1_points)
        l3_xyz, l3_points = sa_forward_ref(model.sa3, l2_xyz, l2_points)
        l4_xyz, l4_points = sa_forward_ref(model.sa4, l3_xyz, l3_points)
        l3_points = fp_forward_ref(model.fp4, l3_xyz, l4_xyz, l3_points, l4_points)
        l2_points = fp_forward_ref(model.fp3, l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = fp_forward_ref(model.fp2, l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = fp_forward_ref(model.fp1, xyz, l1_xyz, points, l1_points)
        x = model.drop1(F.relu(model.bn1(model.conv1(l0_points))))
        x = model.conv2(x)
        return x.permute(0, 2, 1)


def backbone_forward_ref(model, xyz, points):
    """nuScenes backbone `PointNet2.forward` (model/pointmaskrcnn.py:20-32)."""
    with torch.no_grad():
        l1_xyz, l1_points = sa_forward_ref(model.sa1, xyz, points)
        l2_xyz, l2_points = sa_forward_ref(model.sa2, l1_xyz, l1_points)
        l3_xyz, l3_points = sa_forward_ref(model.sa3, l2_xyz, l2_points)
        l4_xyz, l4_points = sa_forward_ref(model.sa4, l3_xyz, l3_points)
        l3_points = fp_forward_ref(model.fp4, l3_xyz, l4_xyz, l3_points, l4_points)
        l2_points = fp_forward_ref(model.fp3, l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = fp_forward_ref(model.fp2, l1_xyz, l2_xyz, l1_points, l2_points)
        return fp_forward_ref(model.fp1, xyz, l1_xyz, None, l1_points)


def multiview_stack_forward_ref(model, xyz, image_features):
    """Point branch of PointNet2Multiview2 / PointNet2Multiview2Msg (model/pointnet2multiview.py:104-120, 216-232)."""
    sa = sa_msg_forward_ref if hasattr(model.sa1_geo, "radius_list") else sa_forward_ref
    with torch.no_grad():
        l1_xyz, l1_points = sa(model.sa1_geo, xyz, None)
        l2_xyz, l2_points = sa(model.sa2_geo, l1_xyz, l1_points)
        l1_xyz_feat, l1_points_feat = sa(model.sa1_feat, xyz, image_features)
        _, l2_points_feat = sa(model.sa2_feat, l1_xyz_feat, l1_points_feat)
        l2_points = torch.cat((l2_points, l2_points_feat), dim=1)
        l3_xyz, l3_points = sa(model.sa3, l2_xyz, l2_points)
        l4_xyz, l4_points = sa(model.sa4, l3_xyz, l3_points)
        l3_points = fp_forward_ref(model.fp4, l3_xyz, l4_xyz, l3_points, l4_points)
        l2_points = fp_forward_ref(model.fp3, l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = fp_forward_ref(model.fp2, l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = fp_forward_ref(model.fp1, xyz, l1_xyz, None, l1_points)
        x = model.drop1(F.relu(model.bn1(model.conv1(l0_points))))
        return model.conv2(x).permute(0, 2, 1)
