"""Evaluation voxelisation on the device: the call surface of the reference's utils/pc_util.py:39-51
(`point_cloud_label_to_surface_voxel_label_fast`) plus the batched form the evaluation loops need.

The reference's loops (train_scannet_semseg.py:204-239, train_scannet_multiview_semseg.py:249-284) bring logits, targets,
weights and coordinates to the host every batch and voxelise scene by scene with numpy.  With `model.predict()` the class
predictions are already one byte per point on the device; `voxel_labels` / `voxel_accuracy_counts` finish the metric
there, so an evaluation step reads back a few dozen counters.
"""
import torch

from . import _lib
from ._lib import ptr


def voxel_first_index(points, mask=None, res=0.0484):
    """points (B, N, >=3) cuda fp32 (only xyz is used), mask (B, N) bool/uint8 or None ->
    uvidx (B, N) fp32, first (B, N) int32 (both padded with -1), count (B,) int32, nvox (B, 3) fp32."""
    _lib.require_cuda(points)
    B, N = points.shape[0], points.shape[1]
    xyz = points[:, :, :3].contiguous().to(torch.float32)
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    dev = points.device
    uvidx = torch.empty((B, N), dtype=torch.float32, device=dev)
    first = torch.empty((B, N), dtype=torch.int32, device=dev)
    count = torch.zeros((B,), dtype=torch.int32, device=dev)
    nvox = torch.empty((B, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("pn2_voxel_first_index", B, N, ptr(xyz), ptr(m), float(res), ptr(uvidx), ptr(first), ptr(count), ptr(nvox),
                  _lib.stream_ptr(dev))
    return uvidx, first, count, nvox


def voxel_labels(points, label, mask=None, res=0.0484):
    """Batched point_cloud_label_to_surface_voxel_label_fast: label (B, N) or (B, N, L) ->
    uvidx (B, N), uvlabel (B, N[, L]) (rows past count[b] are 0), count (B,), nvox (B, 3)."""
    uvidx, first, count, nvox = voxel_first_index(points, mask, res)
    idx = first.clamp_min(0).long()
    valid = first >= 0
    if label.dim() == 2:
        uvlabel = torch.gather(label, 1, idx) * valid.to(label.dtype)
    else:
        uvlabel = torch.gather(label, 1, idx.unsqueeze(-1).expand(-1, -1, label.shape[2])) * valid.unsqueeze(-1).to(label.dtype)
    return uvidx, uvlabel, count, nvox


def point_cloud_label_to_surface_voxel_label_fast(point_cloud, label, res=0.0484):
    """Single-cloud form with the reference's signature and return convention (utils/pc_util.py:39-51) on cuda tensors:
    point_cloud (N, >=3), label (N,) or (N, L) -> (uvidx (U,), uvlabel (U,) or (U, L), nvox (3,))."""
    uvidx, uvlabel, count, nvox = voxel_labels(point_cloud.unsqueeze(0), label.unsqueeze(0), None, res)
    u = int(count[0])
    return uvidx[0, :u], uvlabel[0, :u], nvox[0]


def voxel_accuracy_counts(points, target, pred, weights, num_classes, res=0.02):
    """The voxel-wise counters of the evaluation loop (train_scannet_semseg.py:225-239), summed over the batch, on the
    device: dict of total_correct_vox, total_seen_vox, labelweights_vox (C,), seen/correct/union per class (C,)."""
    mask = weights > 0
    lab = torch.stack((target.long(), pred.long()), dim=-1)
    _, uv, count, _ = voxel_labels(points, lab, mask, res)
    n = uv.shape[1]
    valid = torch.arange(n, device=uv.device).unsqueeze(0) < count.unsqueeze(1)
    t, p = uv[..., 0], uv[..., 1]
    classes = torch.arange(num_classes, device=uv.device).view(1, 1, -1)
    t_is = (t.unsqueeze(-1) == classes) & valid.unsqueeze(-1)
    p_is = (p.unsqueeze(-1) == classes) & valid.unsqueeze(-1)
    return {
        "total_correct_vox": ((t == p) & (t > 0) & valid).sum(),
        "total_seen_vox": ((t > 0) & valid).sum(),
        "labelweights_vox": t_is.sum(dim=(0, 1)),
        "total_seen_class_vox": t_is.sum(dim=(0, 1)),
        "total_correct_class_vox": (t_is & p_is).sum(dim=(0, 1)),
        "total_union_class_vox": (t_is | p_is).sum(dim=(0, 1)),
    }


class EvalCounters:
    """Accumulates the point-wise and voxel-wise confusion counters of the evaluation loop on the device, two kernel
    launches + one voxelisation per batch (pn2_label_counts, pn2_voxel_first_index); `result()` reads 6 x C integers."""

    def __init__(self, num_classes, device, res=0.02):
        self.num_classes, self.res = num_classes, res
        self.point = torch.zeros((3, num_classes), dtype=torch.int64, device=device)
        self.voxel = torch.zeros((3, num_classes), dtype=torch.int64, device=device)

    def update(self, points, target, pred_u8, weights):
        """points (B, N, >=3) fp32, target (B, N) int64, pred_u8 (B, N) uint8 (model.predict), weights (B, N) fp32"""
        B, N = target.shape
        mask = (weights > 0).to(torch.uint8)
        target = target.contiguous()
        pred_u8 = pred_u8.contiguous()
        _, first, count, _ = voxel_first_index(points, mask, self.res)
        dev = target.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.call("pn2_label_counts", B, N, self.num_classes, None, None, ptr(mask), ptr(target), ptr(pred_u8), ptr(self.point), st)
            _lib.call("pn2_label_counts", B, N, self.num_classes, ptr(first), ptr(count), None, ptr(target), ptr(pred_u8), ptr(self.voxel), st)

    def result(self):
        p, v = self.point.cpu().numpy(), self.voxel.cpu().numpy()
        return {"total_seen": int(p[0, 1:].sum()), "total_correct": int(p[1, 1:].sum()),
                "total_seen_class": p[0], "total_correct_class": p[1], "total_union_class": p[2],
                "total_seen_vox": int(v[0, 1:].sum()), "total_correct_vox": int(v[1, 1:].sum()), "labelweights_vox": v[0],
                "total_seen_class_vox": v[0], "total_correct_class_vox": v[1], "total_union_class_vox": v[2]}
