"""Proposal-layer geometry of the nuScenes detector on the device: the call surface of `iou_spheres` and `nms`
(model/pointmaskrcnn.py:233-321) plus a batched NMS.

The reference builds the IoU table with ~20 torch ops and two `nonzero()` host round trips, and runs the greedy
suppression as a Python loop with one synchronisation per kept sphere; here each is one kernel launch (csrc/sphere.cu).
The vote aggregation of the RPN (model/pointmaskrcnn.py:107, 136-139) is an ordinary set-abstraction block
(`PointNetSetAbstraction(128, 4.0, 32, 128 + 3, [128, 128, 128], False)`) and runs on the fused SA path unchanged.
"""
import torch

from . import _lib
from ._lib import ptr


def _spheres(t):
    _lib.require_cuda(t)
    return t.to(torch.float32).contiguous()


def iou_spheres(spheres_a, spheres_b, no_grad=False):
    """spheres_a (M, 4), spheres_b (N, 4) as (x, y, z, r) -> IoU table (M, N).  Forward only (`no_grad` is accepted for
    signature compatibility; the reference's differentiable branch computes the same values)."""
    a, b = _spheres(spheres_a), _spheres(spheres_b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("pn2_sphere_iou", a.shape[0], b.shape[0], ptr(a), ptr(b), ptr(out), _lib.stream_ptr(a.device))
    return out


def nms_batched(bspheres, scores, threshold=0.7, counts=None):
    """bspheres (B, N, 4), scores (B, N), counts (B,) int32 or None -> keep (B, N) int32 (selection order, -1 padded),
    count (B,) int32."""
    s = _spheres(bspheres)
    sc = scores.to(torch.float32).contiguous()
    B, N = sc.shape
    keep = torch.empty((B, N), dtype=torch.int32, device=s.device)
    count = torch.zeros((B,), dtype=torch.int32, device=s.device)
    cin = None if counts is None else counts.to(torch.int32).contiguous()
    with torch.cuda.device(s.device):
        _lib.call("pn2_sphere_nms", B, N, ptr(s), ptr(sc), ptr(cin), float(threshold), ptr(keep), ptr(count), _lib.stream_ptr(s.device))
    return keep, count


def nms(bspheres, scores, threshold=0.7):
    """bspheres (N, 4), scores (N,) -> kept indices, int64, in selection order (model/pointmaskrcnn.py:290-321)."""
    keep, count = nms_batched(bspheres.unsqueeze(0), scores.unsqueeze(0), threshold)
    return keep[0, :int(count[0])].long()
