"""Networks on the hot path, mirroring the reference's callers of the SA/FP modules.

PointNet2SemSeg      model/pointnet2.py:131-162          (ScanNet SSG semseg -- the benchmark model)
PointNet2Backbone    model/pointmaskrcnn.py:8-32         (nuScenes backbone `PointNet2`)

Attribute names match the reference so its state_dicts load unchanged.  In eval mode under
torch.no_grad() the forward is 17 kernel launches (2 layout transposes, 4 x (FPS+gather, ball
query, fused SA), 4 x (three_nn+weights, fused FP; the head is folded into the last FP block) -- the reference issues ~170
(SURVEY.md 3.2).  Otherwise the reference's own composition runs (training / autograd).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .pointnet_util import (PointNetFeaturePropagation, PointNetSetAbstraction, _FoldCache, _fusable, get_mlp_precision,
                            three_nn_weights_cl,
                            to_channel_last)


class PointNet2SemSeg(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.sa1 = PointNetSetAbstraction(1024, 0.1, 32, 3 + 3, [32, 32, 64], False)
        self.sa2 = PointNetSetAbstraction(256, 0.2, 32, 64 + 3, [64, 64, 128], False)
        self.sa3 = PointNetSetAbstraction(64, 0.4, 32, 128 + 3, [128, 128, 256], False)
        self.sa4 = PointNetSetAbstraction(16, 0.8, 32, 256 + 3, [256, 256, 512], False)
        self.fp4 = PointNetFeaturePropagation(768, [256, 256])
        self.fp3 = PointNetFeaturePropagation(384, [256, 256])
        self.fp2 = PointNetFeaturePropagation(320, [256, 128])
        self.fp1 = PointNetFeaturePropagation(128 + 3, [128, 128, 128])
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)
        self._head_fold = _FoldCache()
        self.timers = None  # bench.py: dict name -> [(start_event, end_event)] recorded on the current stream

    @property
    def compute_dtype(self):
        return "bf16" if get_mlp_precision() == "bf16" else "f32"

    def _fp1_with_head(self):
        """fp1's three layers + conv1/bn1/relu (+ eval dropout = identity) + conv2 as ONE fused stack."""
        convs = list(self.fp1.mlp_convs) + [self.conv1, self.conv2]
        bns = list(self.fp1.mlp_bns) + [self.bn1, None]
        relus = [True] * len(self.fp1.mlp_convs) + [True, False]
        return self._head_fold.get(convs, bns, relus)

    def forward_fused(self, xyz, points):
        """xyz (B, 3, N), points (B, D, N) -> (B, N, num_classes), contiguous."""
        xyz_cl, feat_cl = to_channel_last(xyz), to_channel_last(points)
        l1_xyz, l1 = self.sa1.forward_cl(xyz_cl, feat_cl)
        l2_xyz, l2 = self.sa2.forward_cl(l1_xyz, l1)
        l3_xyz, l3 = self.sa3.forward_cl(l2_xyz, l2)
        l4_xyz, l4 = self.sa4.forward_cl(l3_xyz, l3)
        l3 = self.fp4.forward_cl(l3_xyz, l4_xyz, l3, l4)
        l2 = self.fp3.forward_cl(l2_xyz, l3_xyz, l2, l3)
        l1 = self.fp2.forward_cl(l1_xyz, l2_xyz, l1, l2)
        nnw = three_nn_weights_cl(xyz_cl, l1_xyz)
        if self.timers is None:
            return self.fp1.forward_cl(xyz_cl, l1_xyz, feat_cl, l1, mlp=self._fp1_with_head(), nn_weights=nnw)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = self.fp1.forward_cl(xyz_cl, l1_xyz, feat_cl, l1, mlp=self._fp1_with_head(), nn_weights=nnw)
        end.record()
        self.timers.setdefault("fp1_head", []).append((start, end))
        return out

    def forward(self, xyz, points):
        if _fusable(self, xyz, points):
            return self.forward_fused(xyz, points)
        l1_xyz, l1_points = self.sa1(xyz, points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l4_xyz, l4_points = self.sa4(l3_xyz, l3_points)
        l3_points = self.fp4(l3_xyz, l4_xyz, l3_points, l4_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(xyz, l1_xyz, points, l1_points)
        x = self.drop1(F.relu(self.bn1(self.conv1(l0_points))))
        x = self.conv2(x)
        return x.permute(0, 2, 1)


class PointNet2Backbone(nn.Module):
    """The nuScenes backbone (`PointNet2`, model/pointmaskrcnn.py:8-32): xyz (B,3,N), points (B,2,N) ->
    (B, 128, N) per-point features."""

    def __init__(self):
        super().__init__()
        self.sa1 = PointNetSetAbstraction(4096, 1.0, 32, 2 + 3, [32, 32, 64], False)
        self.sa2 = PointNetSetAbstraction(1024, 2.0, 32, 64 + 3, [64, 64, 128], False)
        self.sa3 = PointNetSetAbstraction(256, 4.0, 32, 128 + 3, [128, 128, 256], False)
        self.sa4 = PointNetSetAbstraction(64, 8.0, 32, 256 + 3, [256, 256, 512], False)
        self.fp4 = PointNetFeaturePropagation(768, [256, 256])
        self.fp3 = PointNetFeaturePropagation(384, [256, 256])
        self.fp2 = PointNetFeaturePropagation(320, [256, 128])
        self.fp1 = PointNetFeaturePropagation(128, [128, 128, 128])

    def forward_fused(self, xyz, points):
        """-> (B, N, 128) channel-last"""
        xyz_cl, feat_cl = to_channel_last(xyz), to_channel_last(points)
        l1_xyz, l1 = self.sa1.forward_cl(xyz_cl, feat_cl)
        l2_xyz, l2 = self.sa2.forward_cl(l1_xyz, l1)
        l3_xyz, l3 = self.sa3.forward_cl(l2_xyz, l2)
        l4_xyz, l4 = self.sa4.forward_cl(l3_xyz, l3)
        l3 = self.fp4.forward_cl(l3_xyz, l4_xyz, l3, l4)
        l2 = self.fp3.forward_cl(l2_xyz, l3_xyz, l2, l3)
        l1 = self.fp2.forward_cl(l1_xyz, l2_xyz, l1, l2)
        return self.fp1.forward_cl(xyz_cl, l1_xyz, None, l1)

    def forward(self, xyz, points):
        if _fusable(self, xyz, points):
            return self.forward_fused(xyz, points).permute(0, 2, 1)
        l1_xyz, l1_points = self.sa1(xyz, points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l4_xyz, l4_points = self.sa4(l3_xyz, l3_points)
        l3_points = self.fp4(l3_xyz, l4_xyz, l3_points, l4_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        return self.fp1(xyz, l1_xyz, None, l1_points)
