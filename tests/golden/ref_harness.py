"""Runs the REFERENCE's own, unmodified Python (model/pointnet_util.py, model/pointnet2.py, model/pointnet2multiview.py,
utils/projection.py) on the CPU of the build container, so that golden fixtures come from the reference itself and not
from a restatement (TEST INFRASTRUCTURE ONLY; used by the make_golden_*.py generators, never by the tests or the product).

What is substituted, and only this:
  * `pointnet2_cuda` (utils/src/pointnet2_api.cpp:10-23 -- the CUDA extension this repo replaces) is a module whose nine
    functions run the C oracle (oracle/pn2_oracle.c, itself pinned bit-exact to the reference's compiled kernels by
    tests/golden/ref_cuda_r1.npz) on host tensors;
  * `torch.cuda.IntTensor/FloatTensor` (model/pointnet2_utils.py:25-26,55,...) and `Tensor.cuda()` (utils/projection.py:108,
    181,191,215) resolve to the host;
  * `model.enet.create_enet_for_3d` (needs ./scannetv2_enet.pth, not shipped) returns identities: the 2-D CNN is out of scope,
    its OUTPUT feature maps are the inputs of the lifting.
Everything else -- every torch op, its order, the module classes, their state_dict layout -- is the reference's code.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
for p in (ROOT, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import oracle as orc  # noqa: E402


def _np(t):
    return np.ascontiguousarray(t.detach().numpy())


def _put(dst, arr):
    dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).reshape(dst.shape))


def oracle_backed_pointnet2_cuda():
    """A module with the nine pybind names and argument orders of utils/src/pointnet2_api.cpp:10-23, running on the host."""
    m = types.ModuleType("pointnet2_cuda")

    def furthest_point_sampling_wrapper(b, n, npoint, xyz, temp, out):
        _put(out, orc.furthest_point_sample(_np(xyz), npoint))
        return 1

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        _put(out, orc.gather_operation(_np(points), _np(idx)))
        return 1

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        _put(grad_points, orc.gather_operation_grad(_np(grad_out), _np(idx), n))
        return 1

    def ball_query_wrapper(b, n, m_, radius, nsample, new_xyz, xyz, idx):
        _put(idx, orc.ball_query(radius, nsample, _np(xyz), _np(new_xyz)))
        return 1

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        _put(out, orc.grouping_operation(_np(points), _np(idx)))
        return 1

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        _put(grad_points, orc.grouping_operation_grad(_np(grad_out), _np(idx), n))
        return 1

    def three_nn_wrapper(b, n, m_, unknown, known, dist2, idx):
        d2, i = orc.three_nn_dist2(_np(unknown), _np(known))
        _put(dist2, d2)
        _put(idx, i)

    def three_interpolate_wrapper(b, c, m_, n, points, idx, weight, out):
        _put(out, orc.three_interpolate(_np(points), _np(idx), _np(weight)))

    def three_interpolate_grad_wrapper(b, c, n, m_, grad_out, idx, weight, grad_points):
        _put(grad_points, orc.three_interpolate_grad(_np(grad_out), _np(idx), _np(weight), m_))

    for f in (furthest_point_sampling_wrapper, gather_points_wrapper, gather_points_grad_wrapper, ball_query_wrapper,
              group_points_wrapper, group_points_grad_wrapper, three_nn_wrapper, three_interpolate_wrapper,
              three_interpolate_grad_wrapper):
        setattr(m, f.__name__, f)
    return m


@contextlib.contextmanager
def reference_on_cpu():
    """Context in which `import model.pointnet2`, `import utils.projection` ... give the reference's modules, host-only."""
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("the reference tree is needed to generate fixtures (it is absent on the GPU box by design)")
    saved_mods = {k: sys.modules.get(k) for k in ("pointnet2_cuda", "model", "utils")}
    saved = (torch.Tensor.cuda, torch.cuda.IntTensor, torch.cuda.FloatTensor, torch.nn.Module.cuda)
    sys.modules["pointnet2_cuda"] = oracle_backed_pointnet2_cuda()
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.IntTensor = torch.IntTensor
    torch.cuda.FloatTensor = torch.FloatTensor
    try:
        import model.enet as enet
        enet.create_enet_for_3d = lambda *a, **k: (torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity())
        yield
    finally:
        torch.Tensor.cuda, torch.cuda.IntTensor, torch.cuda.FloatTensor, torch.nn.Module.cuda = saved
        sys.path.remove(REFERENCE)
        for k in list(sys.modules):
            if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
