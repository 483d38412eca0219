"""Loader for oracle/_ref/ref_cuda.so: the reference's OWN four .cu files (utils/src/*_gpu.cu) compiled
unmodified for sm_100a by oracle/build.py --ref, behind our extern "C" shim (oracle/ref_shim.cu).

TEST INFRASTRUCTURE ONLY (tests/, bench.py's reference legs).  The Python side below reproduces the
caller-side conventions of model/pointnet2_utils.py: temp filled with 1e10 (:26), idx zero-filled for
ball_query (:216), sqrt of three_nn's squared distances (:97), zeroed gradient buffers (:67,144,188).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "ref_cuda.so")
_lib = None


def available():
    return os.path.exists(PATH) and torch.cuda.is_available()


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
    return _lib


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def furthest_point_sample(xyz, npoint):
    B, N, _ = xyz.shape
    out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    lib().ref_fps(B, N, npoint, _p(xyz), _p(temp), _p(out), _s())
    return out


def gather_operation(features, idx):
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=features.device)
    lib().ref_gather(B, C, N, M, _p(features), _p(idx), _p(out), _s())
    return out


def gather_operation_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    lib().ref_gather_grad(B, C, N, M, _p(grad_out), _p(idx), _p(g), _s())
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)
    lib().ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def grouping_operation(features, idx):
    B, C, N = features.shape
    _, P, S = idx.shape
    out = torch.empty((B, C, P, S), dtype=torch.float32, device=features.device)
    lib().ref_group(B, C, N, P, S, _p(features), _p(idx), _p(out), _s())
    return out


def grouping_operation_grad(grad_out, idx, N):
    B, C, P, S = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    lib().ref_group_grad(B, C, N, P, S, _p(grad_out), _p(idx), _p(g), _s())
    return g


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    lib().ref_three_nn(B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return torch.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    B, C, m = features.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=features.device)
    lib().ref_three_interpolate(B, C, m, n, _p(features), _p(idx), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    g = torch.zeros((B, C, m), dtype=torch.float32, device=grad_out.device)
    lib().ref_three_interpolate_grad(B, C, n, m, _p(grad_out), _p(idx), _p(weight), _p(g), _s())
    return g
