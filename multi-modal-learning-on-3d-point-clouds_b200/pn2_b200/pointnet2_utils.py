"""Operator layer: the call surface of the reference's model/pointnet2_utils.py on the B200 kernels.

Same names, argument order, shapes, dtypes and semantics as the reference:
    furthest_point_sample (:10-36)   gather_operation (:39-73)   three_nn (:76-104)
    three_interpolate (:107-151)     grouping_operation (:154-195)   ball_query (:198-226)
    QueryAndGroup (:229-262)         GroupAll (:265-288)
plus the four names BASELINE.json's north_star lists (farthest_point_sample, query_ball_point,
index_points, square_distance) as thin aliases with the REFERENCE's semantics.

All operators require contiguous CUDA tensors and raise otherwise -- there is no CPU path.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import pointnet2_cuda as _ext
from ._lib import Pn2Error, require_cuda


_DETERMINISTIC = None  # None: follow torch.are_deterministic_algorithms_enabled()


def set_deterministic(flag):
    """True: the gather / grouping / three_interpolate backwards sum every gradient element in a fixed order (inverse
    index + segmented reduction, pn2_scatter_rows_det) instead of the reference's atomicAdd scatter, so training runs are
    bit-reproducible.  None (default): follow torch.use_deterministic_algorithms().  Returns the previous setting."""
    global _DETERMINISTIC
    prev, _DETERMINISTIC = _DETERMINISTIC, flag
    return prev


def _deterministic():
    return torch.are_deterministic_algorithms_enabled() if _DETERMINISTIC is None else bool(_DETERMINISTIC)


def _scatter_det(grad_out, idx, n, weight=None):
    """grad_out (B, C, ...) contiguous, idx (B, ...) int32 with values in [0, n) -> (B, C, n); see include/pn2_abi.h"""
    from . import _lib
    B, C = grad_out.shape[0], grad_out.shape[1]
    J = idx[0].numel()
    dev = grad_out.device
    seg = torch.empty((B * n + 1,), dtype=torch.int32, device=dev)
    pos = torch.empty((B, max(J, 1)), dtype=torch.int32, device=dev)
    grad = torch.zeros((B, C, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.call("pn2_inverse_index", B, n, J, _lib.ptr(idx), _lib.ptr(seg), _lib.ptr(pos), st)
        _lib.call("pn2_scatter_rows_det", B, C, n, J, 3 if weight is not None else 1, _lib.ptr(grad_out), _lib.ptr(seg), _lib.ptr(pos),
                  _lib.ptr(weight), _lib.ptr(grad), st)
    return grad


def _grid_pays_off(n_points, n_queries):
    """Cell-list search (one sort per cloud) instead of brute force: worthwhile for mid-sized clouds with many queries."""
    from .pointnet_util import grid_max_points
    return 512 <= n_points <= grid_max_points() and n_points * n_queries >= (1 << 20)


def _need_contiguous(**tensors):
    for name, t in tensors.items():
        if not t.is_contiguous():
            raise Pn2Error("%s must be contiguous (the reference asserts the same)" % name)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B, N, 3) fp32 -> (B, npoint) int32; index 0 is always picked first."""
        require_cuda(xyz)
        _need_contiguous(xyz=xyz)
        B, N, _ = xyz.size()
        out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        # the reference allocates a (B, N) scratch of 1e10 here (:26); running minima live on chip
        # in our kernels, so the scratch is only materialised for clouds beyond their capacity.
        temp = torch.empty((B, N), dtype=torch.float32, device=xyz.device) if N > 49152 else None  # beyond the on-chip kernels
        _ext.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint) int32 -> (B, C, npoint)"""
        require_cuda(features, idx)
        _need_contiguous(features=features, idx=idx)
        B, npoint = idx.size()
        _, C, N = features.size()
        out = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        _ext.gather_points_wrapper(B, C, N, npoint, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.dims = (C, N)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        C, N = ctx.dims
        B, npoint = idx.size()
        if _deterministic():
            return _scatter_det(grad_out.contiguous(), idx, N), None
        grad = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _ext.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad)
        return grad, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B, n, 3), known (B, m, 3) -> (dist (B, n, 3) L2 distances, idx (B, n, 3) int32)"""
        require_cuda(unknown, known)
        _need_contiguous(unknown=unknown, known=known)
        B, n, _ = unknown.size()
        m = known.size(1)
        if _grid_pays_off(m, n) and m >= 512:
            from .pointnet_util import SpatialGrid
            idx, dist2 = SpatialGrid(known, 0.0).three_nn(unknown, want_dist2=True, want_weight=False)
        else:
            dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
            idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
            _ext.three_nn_wrapper(B, n, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)  # the kernel returns squared distances (:97)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B, C, m), idx (B, n, 3) int32, weight (B, n, 3) -> (B, C, n)"""
        require_cuda(features, idx, weight)
        _need_contiguous(features=features, idx=idx, weight=weight)
        B, C, m = features.size()
        n = idx.size(1)
        out = torch.empty((B, C, n), dtype=torch.float32, device=features.device)
        _ext.three_interpolate_wrapper(B, C, m, n, features, idx, weight, out)
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        B, C, n = grad_out.size()
        if _deterministic():
            return _scatter_det(grad_out.contiguous(), idx, ctx.m, weight=weight), None, None
        grad = torch.zeros((B, C, ctx.m), dtype=torch.float32, device=grad_out.device)
        _ext.three_interpolate_grad_wrapper(B, C, n, ctx.m, grad_out.contiguous(), idx, weight, grad)
        return grad, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint, nsample) int32 -> (B, C, npoint, nsample)"""
        require_cuda(features, idx)
        _need_contiguous(features=features, idx=idx)
        B, npoint, nsample = idx.size()
        _, C, N = features.size()
        out = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=features.device)
        _ext.group_points_wrapper(B, C, N, npoint, nsample, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, npoint, nsample = grad_out.size()
        if _deterministic():
            return _scatter_det(grad_out.contiguous(), idx, ctx.N), None
        grad = torch.zeros((B, C, ctx.N), dtype=torch.float32, device=grad_out.device)
        _ext.group_points_grad_wrapper(B, C, ctx.N, npoint, nsample, grad_out.contiguous(), idx, grad)
        return grad, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B, N, 3), new_xyz (B, npoint, 3) -> idx (B, npoint, nsample) int32"""
        require_cuda(xyz, new_xyz)
        _need_contiguous(xyz=xyz, new_xyz=new_xyz)
        B, N, _ = xyz.size()
        if _grid_pays_off(N, new_xyz.size(1)):
            # exact same result through a cell list instead of the brute-force scan (csrc/grid.cu)
            from .pointnet_util import SpatialGrid
            idx = SpatialGrid(xyz, 1.01 * float(radius)).ball_query(radius, nsample, new_xyz)
        else:
            idx = BallQuery.brute_force(radius, nsample, xyz, new_xyz)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def brute_force(radius, nsample, xyz, new_xyz):
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)  # kernel writes every slot
        _ext.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """ball_query + grouping + centring, xyz channels first (reference :229-262)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None) -> torch.Tensor:
        """xyz (B, N, 3), new_xyz (B, npoint, 3), features (B, C, N) -> (B, 3 + C, npoint, nsample)"""
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            if not self.use_xyz:
                raise AssertionError("Cannot have not features and not use xyz as a feature!")
            return grouped_xyz
        grouped = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped


class GroupAll(nn.Module):
    """One group holding every point (reference :265-288)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None) -> torch.Tensor:
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped


# ---- north_star aliases, reference semantics (SURVEY.md 8a "name map") -----------------------------

def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """xyz (B, N, 3) -> (B, npoint) int32.  Starts at index 0 (not random) with the reference's tie order."""
    return furthest_point_sample(xyz.contiguous(), npoint)


def query_ball_point(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """-> (B, S, nsample) int32; strict '<' on the direct-difference fp32 distance."""
    return ball_query(radius, nsample, xyz.contiguous(), new_xyz.contiguous())


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points (B, N, C) channel-last, idx (B, S) or (B, S, K) -> (B, S, C) or (B, S, K, C)."""
    feats = points.transpose(1, 2).contiguous()
    idx = idx.to(torch.int32).contiguous()
    if idx.dim() == 2:
        return gather_operation(feats, idx).transpose(1, 2).contiguous()
    return grouping_operation(feats, idx).permute(0, 2, 3, 1).contiguous()


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """src (B, N, 3), dst (B, M, 3) -> (B, N, M) with the reference kernels' rounding sequence
    fma(dz,dz, fma(dx,dx, rn(dy*dy))) -- never the -2ab+a^2+b^2 matmul expansion.
    Elementwise torch ops (addcmul is a fused multiply-add on CUDA); used for inspection, not on the hot path."""
    d = src.unsqueeze(2) - dst.unsqueeze(1)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    t = dy * dy
    t = torch.addcmul(t, dx, dx)
    return torch.addcmul(t, dz, dz)
