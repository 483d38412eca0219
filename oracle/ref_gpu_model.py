"""The reference's REAL implementation of the path, on the GPU: its own CUDA kernels (oracle/_ref/ref_cuda.so,
compiled verbatim) driven through the same operator sequence model/pointnet_util.py issues -- every
transposing copy included -- with stock torch.nn Conv/BatchNorm (cuDNN/cuBLAS) for the shared MLPs.

TEST / BENCH INFRASTRUCTURE ONLY.  bench.py reports its throughput next to ours as `ref_gpu`.  The
reference's Python files cannot travel to the GPU box and its pybind glue no longer compiles against
torch 2.x, so the operator sequence is restated here (model/pointnet_util.py:20-47, 85-111, 185-221;
model/pointnet2.py:147-162): what matters for the timing is that the same kernels, the same layout
copies and the same cuDNN calls run in the same order.
"""
import torch
import torch.nn.functional as F

from . import ref_cuda as R


def _swap(t):
    """(B, X, Y) -> contiguous (B, Y, X): one transposing copy kernel, as each .permute().contiguous() in the reference"""
    return t.transpose(1, 2).contiguous()


def _shared_mlp(x, convs, bns):
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return x


def _group(level, pts_cl, feat_cl):
    """sample_and_group: FPS -> gather -> ball query -> group xyz / features -> centre -> concat (xyz first)."""
    B, _, C = pts_cl.shape
    pts_cf = _swap(pts_cl)
    picks = R.furthest_point_sample(pts_cl.contiguous(), level.npoint)
    centres = _swap(R.gather_operation(pts_cf, picks))
    ball = R.ball_query(level.radius, level.nsample, pts_cl.contiguous(), centres.contiguous())
    local = R.grouping_operation(_swap(pts_cl), ball).permute(0, 2, 3, 1).contiguous() - centres.view(B, level.npoint, 1, C)
    if feat_cl is None:
        return centres, local
    gathered = R.grouping_operation(_swap(feat_cl), ball).permute(0, 2, 3, 1).contiguous()
    return centres, torch.cat([local, gathered], dim=-1)


def sa_forward(level, xyz, feats):
    centres, grouped = _group(level, xyz.transpose(1, 2), None if feats is None else feats.transpose(1, 2))
    act = _shared_mlp(grouped.permute(0, 3, 2, 1), level.mlp_convs, level.mlp_bns)
    return centres.transpose(1, 2), act.max(dim=2)[0]


def fp_forward(level, xyz_fine, xyz_coarse, skip, coarse_feats):
    fine, coarse = _swap(xyz_fine), _swap(xyz_coarse)
    B, N, _ = fine.shape
    if coarse.shape[1] == 1:
        up = coarse_feats.transpose(1, 2).repeat(1, N, 1)
    else:
        dist, nn_idx = R.three_nn(fine, coarse)
        dist[dist < 1e-10] = 1e-10
        w = 1.0 / dist
        w = w / w.sum(dim=-1).view(B, N, 1)
        up = _swap(R.three_interpolate(_swap(coarse_feats.transpose(1, 2)), nn_idx, w))
    rows = up if skip is None else torch.cat([skip.transpose(1, 2), up], dim=-1)
    return _shared_mlp(rows.transpose(1, 2), level.mlp_convs, level.mlp_bns)


def semseg_forward(net, xyz, feats):
    with torch.no_grad():
        x1, f1 = sa_forward(net.sa1, xyz, feats)
        x2, f2 = sa_forward(net.sa2, x1, f1)
        x3, f3 = sa_forward(net.sa3, x2, f2)
        x4, f4 = sa_forward(net.sa4, x3, f3)
        f3 = fp_forward(net.fp4, x3, x4, f3, f4)
        f2 = fp_forward(net.fp3, x2, x3, f2, f3)
        f1 = fp_forward(net.fp2, x1, x2, f1, f2)
        f0 = fp_forward(net.fp1, xyz, x1, feats, f1)
        logits = net.conv2(net.drop1(F.relu(net.bn1(net.conv1(f0)))))
        return logits.transpose(1, 2)
