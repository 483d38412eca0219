"""Drop-in for the reference's compiled extension module `pointnet2_cuda`
(utils/src/pointnet2_api.cpp:10-23): the same nine functions, positional arguments in the same
order (ints first, tensors last, outputs caller-allocated and written in place), launched
asynchronously on the current stream of the tensors' device.

Differences by design: failures raise `Pn2Error` instead of `exit(-1)`
(utils/src/sampling_gpu.cu:39-43), and inputs are validated (the reference checks only ball_query).
"""
import torch

from . import _lib


def _check(name, t, dtype, numel=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.Pn2Error("%s must be a CUDA tensor (no CPU path)" % name)
    if t.dtype != dtype:
        raise _lib.Pn2Error("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.Pn2Error("%s must be contiguous" % name)
    if numel is not None and t.numel() < numel:
        raise _lib.Pn2Error("%s has %d elements, needs %d" % (name, t.numel(), numel))


def _run(name, dev, *args):
    with torch.cuda.device(dev):
        _lib.call(name, *args, _lib.stream_ptr(dev))


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _check("new_xyz", new_xyz, torch.float32, b * m * 3)
    _check("xyz", xyz, torch.float32, b * n * 3)
    _check("idx", idx, torch.int32, b * m * nsample)
    _run("pn2_ball_query", xyz.device, b, n, m, float(radius), nsample, _lib.ptr(new_xyz), _lib.ptr(xyz), _lib.ptr(idx))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _check("points", points, torch.float32, b * c * n)
    _check("idx", idx, torch.int32, b * npoints * nsample)
    _check("out", out, torch.float32, b * c * npoints * nsample)
    _run("pn2_group_points", points.device, b, c, n, npoints, nsample, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(out))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _check("grad_out", grad_out, torch.float32, b * c * npoints * nsample)
    _check("idx", idx, torch.int32, b * npoints * nsample)
    _check("grad_points", grad_points, torch.float32, b * c * n)
    _run("pn2_group_points_grad", grad_out.device, b, c, n, npoints, nsample, _lib.ptr(grad_out), _lib.ptr(idx),
         _lib.ptr(grad_points))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _check("points", points, torch.float32, b * c * n)
    _check("idx", idx, torch.int32, b * npoints)
    _check("out", out, torch.float32, b * c * npoints)
    _run("pn2_gather_points", points.device, b, c, n, npoints, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(out))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _check("grad_out", grad_out, torch.float32, b * c * npoints)
    _check("idx", idx, torch.int32, b * npoints)
    _check("grad_points", grad_points, torch.float32, b * c * n)
    _run("pn2_gather_points_grad", grad_out.device, b, c, n, npoints, _lib.ptr(grad_out), _lib.ptr(idx),
         _lib.ptr(grad_points))
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _check("points", points, torch.float32, b * n * 3)
    _check("idx", idx, torch.int32, b * m)
    if temp is not None:
        _check("temp", temp, torch.float32, b * n)
    _run("pn2_furthest_point_sampling", points.device, b, n, m, _lib.ptr(points), _lib.ptr(temp), _lib.ptr(idx))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _check("unknown", unknown, torch.float32, b * n * 3)
    _check("known", known, torch.float32, b * m * 3)
    _check("dist2", dist2, torch.float32, b * n * 3)
    _check("idx", idx, torch.int32, b * n * 3)
    _run("pn2_three_nn", unknown.device, b, n, m, _lib.ptr(unknown), _lib.ptr(known), _lib.ptr(dist2), _lib.ptr(idx))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _check("points", points, torch.float32, b * c * m)
    _check("idx", idx, torch.int32, b * n * 3)
    _check("weight", weight, torch.float32, b * n * 3)
    _check("out", out, torch.float32, b * c * n)
    _run("pn2_three_interpolate", points.device, b, c, m, n, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(weight), _lib.ptr(out))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _check("grad_out", grad_out, torch.float32, b * c * n)
    _check("idx", idx, torch.int32, b * n * 3)
    _check("weight", weight, torch.float32, b * n * 3)
    _check("grad_points", grad_points, torch.float32, b * c * m)
    _run("pn2_three_interpolate_grad", grad_out.device, b, c, n, m, _lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(weight),
         _lib.ptr(grad_points))
