"""Differentiable CPU restatement of one SA + one FP block in TRAINING mode (TEST INFRASTRUCTURE ONLY).

Indices (FPS, ball query, 3-NN) come from the C oracle exactly as the reference obtains them from its
non-differentiable kernels (model/pointnet2_utils.py:32-33, 101-102, 223-224 return None gradients); the gathers and the
interpolation are written with torch indexing so autograd provides the reference gradients of
gather_points_grad / group_points_grad / three_interpolate_grad (utils/src/*_gpu.cu backward kernels: scatter-adds)."""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _idx(a):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int64))


def sa_train_ref(mod, xyz, points):
    """PointNetSetAbstraction.forward (model/pointnet_util.py:85-111) with autograd through the grouping."""
    B, _, N = xyz.shape
    xyz_np = np.ascontiguousarray(xyz.detach().permute(0, 2, 1).numpy(), dtype=np.float32)
    fps = orc.furthest_point_sample(xyz_np, mod.npoint)
    new_xyz_np = np.take_along_axis(xyz_np, fps[..., None].astype(np.int64), 1)
    ball = _idx(orc.ball_query(mod.radius, mod.nsample, xyz_np, new_xyz_np))          # (B, S, K)
    S, K = mod.npoint, mod.nsample
    flat = ball.reshape(B, 1, S * K)
    grouped_xyz = torch.gather(xyz, 2, flat.expand(B, 3, S * K)).reshape(B, 3, S, K)
    new_xyz = torch.from_numpy(new_xyz_np).permute(0, 2, 1)                            # (B, 3, S)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(-1)
    if points is not None:
        D = points.shape[1]
        grouped_pts = torch.gather(points, 2, flat.expand(B, D, S * K)).reshape(B, D, S, K)
        feats = torch.cat([grouped_xyz, grouped_pts], dim=1)                           # xyz first (:41)
    else:
        feats = grouped_xyz
    feats = feats.permute(0, 1, 3, 2)                                                  # (B, C, K, S)
    for conv, bn in zip(mod.mlp_convs, mod.mlp_bns):
        feats = F.relu(bn(conv(feats)))
    return new_xyz, torch.max(feats, 2)[0]


def fp_train_ref(mod, xyz1, xyz2, points1, points2):
    """PointNetFeaturePropagation.forward (model/pointnet_util.py:185-221) with autograd through the interpolation."""
    B, _, N = xyz1.shape
    x1 = np.ascontiguousarray(xyz1.detach().permute(0, 2, 1).numpy(), dtype=np.float32)
    x2 = np.ascontiguousarray(xyz2.detach().permute(0, 2, 1).numpy(), dtype=np.float32)
    dist, idx = orc.three_nn(x1, x2)
    w = torch.from_numpy(orc.fp_weights(dist))                                         # (B, N, 3), constants as in the reference
    idx = _idx(idx)
    D2 = points2.shape[1]
    g = torch.gather(points2, 2, idx.reshape(B, 1, N * 3).expand(B, D2, N * 3)).reshape(B, D2, N, 3)
    interp = (g * w.unsqueeze(1)).sum(-1)
    new_points = torch.cat([points1, interp], dim=1) if points1 is not None else interp
    for conv, bn in zip(mod.mlp_convs, mod.mlp_bns):
        new_points = F.relu(bn(conv(new_points)))
    return new_points
