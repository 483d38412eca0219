"""`torch.library` registration of the path's operators: torch.ops.pn2.* (SURVEY.md 7 step 2 / 8b).

The reference exposes its kernels through `torch.autograd.Function` subclasses (model/pointnet2_utils.py:10-226); those
classes stay the drop-in surface (pn2_b200/pointnet2_utils.py).  This module registers the SAME kernels as dispatcher-level
custom operators, with
  * a CUDA implementation that calls the C ABI through the `pointnet2_cuda` shim (no CPU kernel is registered: a CPU tensor
    raises the dispatcher's "no kernel for CPU" error -- there is no fallback),
  * a fake (meta) kernel giving output shapes / dtypes, so FakeTensor propagation, `torch.export` and `torch.compile`
    tracing see through them instead of graph-breaking on ctypes calls,
  * autograd formulas whose backwards are themselves registered ops (pn2::*_grad), i.e. the scatter-adds of
    utils/src/{sampling,group_points,interpolate}_gpu.cu (K3, K6, K9), or the deterministic segmented sums when
    deterministic algorithms are requested.

    import pn2_b200.ops                       # registers the library once
    idx = torch.ops.pn2.furthest_point_sample(xyz, 1024)

Operator schemas (all index tensors int32, all float tensors contiguous fp32, as the reference asserts):
    furthest_point_sample(Tensor xyz, int npoint) -> Tensor                     (B, npoint) int32
    gather_operation(Tensor features, Tensor idx) -> Tensor                     (B, C, npoint)
    three_nn(Tensor unknown, Tensor known) -> (Tensor dist, Tensor idx)         (B, n, 3), (B, n, 3) int32
    three_interpolate(Tensor features, Tensor idx, Tensor weight) -> Tensor     (B, C, n)
    grouping_operation(Tensor features, Tensor idx) -> Tensor                   (B, C, npoint, nsample)
    ball_query(float radius, int nsample, Tensor xyz, Tensor new_xyz) -> Tensor (B, npoint, nsample) int32
"""
from typing import Tuple

import torch

from . import pointnet2_utils as _pu

_LIB = "pn2"


def _f32(t, name):
    if t.dtype != torch.float32:
        raise TypeError("pn2::%s expects float32, got %s" % (name, t.dtype))


def _i32(t, name):
    if t.dtype != torch.int32:
        raise TypeError("pn2::%s expects an int32 index tensor, got %s" % (name, t.dtype))


# ---- forward operators ------------------------------------------------------------------------------------------------
@torch.library.custom_op(_LIB + "::furthest_point_sample", mutates_args=(), device_types="cuda")
def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    _f32(xyz, "furthest_point_sample")
    return _pu.FurthestPointSampling.apply(xyz, npoint)


@furthest_point_sample.register_fake
def _(xyz, npoint):
    return xyz.new_empty((xyz.shape[0], npoint), dtype=torch.int32)


@torch.library.custom_op(_LIB + "::gather_operation", mutates_args=(), device_types="cuda")
def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _f32(features, "gather_operation")
    _i32(idx, "gather_operation")
    return _pu.GatherOperation.apply(features.detach(), idx)


@gather_operation.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1]))


@torch.library.custom_op(_LIB + "::gather_operation_grad", mutates_args=(), device_types="cuda")
def gather_operation_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    B, C, npoint = grad_out.shape
    grad_out = grad_out.contiguous()
    if _pu._deterministic():
        return _pu._scatter_det(grad_out, idx, n)
    grad = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
    _pu._ext.gather_points_grad_wrapper(B, C, n, npoint, grad_out, idx, grad)
    return grad


@gather_operation_grad.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


def _gather_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _gather_backward(ctx, grad_out):
    (idx,) = ctx.saved_tensors
    return torch.ops.pn2.gather_operation_grad(grad_out, idx, ctx.n), None


gather_operation.register_autograd(_gather_backward, setup_context=_gather_setup)


@torch.library.custom_op(_LIB + "::three_nn", mutates_args=(), device_types="cuda")
def three_nn(unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _f32(unknown, "three_nn")
    _f32(known, "three_nn")
    dist, idx = _pu.ThreeNN.apply(unknown, known)
    return dist, idx


@three_nn.register_fake
def _(unknown, known):
    shape = (unknown.shape[0], unknown.shape[1], 3)
    return unknown.new_empty(shape), unknown.new_empty(shape, dtype=torch.int32)


@torch.library.custom_op(_LIB + "::three_interpolate", mutates_args=(), device_types="cuda")
def three_interpolate(features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    _f32(features, "three_interpolate")
    _f32(weight, "three_interpolate")
    _i32(idx, "three_interpolate")
    return _pu.ThreeInterpolate.apply(features.detach(), idx, weight)


@three_interpolate.register_fake
def _(features, idx, weight):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1]))


@torch.library.custom_op(_LIB + "::three_interpolate_grad", mutates_args=(), device_types="cuda")
def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, m: int) -> torch.Tensor:
    B, C, n = grad_out.shape
    grad_out = grad_out.contiguous()
    if _pu._deterministic():
        return _pu._scatter_det(grad_out, idx, m, weight=weight)
    grad = torch.zeros((B, C, m), dtype=torch.float32, device=grad_out.device)
    _pu._ext.three_interpolate_grad_wrapper(B, C, n, m, grad_out, idx, weight, grad)
    return grad


@three_interpolate_grad.register_fake
def _(grad_out, idx, weight, m):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], m))


def _interp_setup(ctx, inputs, output):
    features, idx, weight = inputs
    ctx.save_for_backward(idx, weight)
    ctx.m = features.shape[2]


def _interp_backward(ctx, grad_out):
    idx, weight = ctx.saved_tensors
    # as in the reference (model/pointnet2_utils.py:133-151) only the features receive a gradient
    return torch.ops.pn2.three_interpolate_grad(grad_out, idx, weight, ctx.m), None, None


three_interpolate.register_autograd(_interp_backward, setup_context=_interp_setup)


@torch.library.custom_op(_LIB + "::grouping_operation", mutates_args=(), device_types="cuda")
def grouping_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _f32(features, "grouping_operation")
    _i32(idx, "grouping_operation")
    return _pu.GroupingOperation.apply(features.detach(), idx)


@grouping_operation.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1], idx.shape[2]))


@torch.library.custom_op(_LIB + "::grouping_operation_grad", mutates_args=(), device_types="cuda")
def grouping_operation_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    B, C, npoint, nsample = grad_out.shape
    grad_out = grad_out.contiguous()
    if _pu._deterministic():
        return _pu._scatter_det(grad_out, idx, n)
    grad = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
    _pu._ext.group_points_grad_wrapper(B, C, n, npoint, nsample, grad_out, idx, grad)
    return grad


@grouping_operation_grad.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


def _group_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _group_backward(ctx, grad_out):
    (idx,) = ctx.saved_tensors
    return torch.ops.pn2.grouping_operation_grad(grad_out, idx, ctx.n), None


grouping_operation.register_autograd(_group_backward, setup_context=_group_setup)


@torch.library.custom_op(_LIB + "::ball_query", mutates_args=(), device_types="cuda")
def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    _f32(xyz, "ball_query")
    _f32(new_xyz, "ball_query")
    return _pu.BallQuery.apply(radius, nsample, xyz, new_xyz)


@ball_query.register_fake
def _(radius, nsample, xyz, new_xyz):
    return xyz.new_empty((xyz.shape[0], new_xyz.shape[1], nsample), dtype=torch.int32)


OPS = ("furthest_point_sample", "gather_operation", "gather_operation_grad", "three_nn", "three_interpolate",
       "three_interpolate_grad", "grouping_operation", "grouping_operation_grad", "ball_query")
