"""Shared by the CPU and GPU lifting parity tests: the committed reference vectors (tests/golden/projection_r2.npz, made by
tests/golden/make_golden_projection.py from the reference's own utils/projection.py) and the classification of a
disagreeing decision as a rounding-boundary case."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROJ = os.path.join(HERE, "golden", "projection_r2.npz")
_spec = importlib.util.spec_from_file_location("make_golden_projection", os.path.join(HERE, "golden", "make_golden_projection.py"))
mgj = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mgj)


def ref_view_params(gold, name):
    """The reference's own per-view parameters in the layout of pn2_lift_views / orc_lift_views:
    w2c (B,V,16), corner2 (B,V,3), corner4 (B,V,3), normals (B,V,18)."""
    cor, nrm, w2c = gold[name + "/corners"], gold[name + "/normals"], gold[name + "/w2c"]
    B, V = cor.shape[:2]
    return (np.ascontiguousarray(w2c.reshape(B, V, 16)), np.ascontiguousarray(cor[:, :, 2, :3]),
            np.ascontiguousarray(cor[:, :, 4, :3]), np.ascontiguousarray(nrm.reshape(B, V, 18)))


def boundary_cases(xyz, depth, w2c, corner2, corner4, normals, where, intr=None, dmin=None, dmax=None, acc=None):
    """where: (n, 3) rows (b, v, i) of decisions that differ between two evaluations of utils/projection.py:166-230.
    -> bool (n,): True when, evaluated in fp64, the point sits within fp32 rounding noise of one of the path's decision
    boundaries: a frustum plane value at -0.005 (round(100 s) flips, :115-117), a pixel coordinate at .5 (:204), the image
    border (:207), the depth limits or |depth - z| = accuracy (:216).  Two correct fp32 evaluations (different summation
    order) can only differ there."""
    intr = intr if intr is not None else (mgj.INTRINSIC[0][0], mgj.INTRINSIC[1][1], mgj.INTRINSIC[0][2], mgj.INTRINSIC[1][2])
    dmin = mgj.DEPTH_MIN if dmin is None else dmin
    dmax = mgj.DEPTH_MAX if dmax is None else dmax
    acc = mgj.ACCURACY if acc is None else acc
    fx, fy, cx, cy = [float(t) for t in intr]
    H, W = depth.shape[-2:]
    ok = np.zeros(len(where), bool)
    for r, (b, v, i) in enumerate(where):
        p = xyz[b, i].astype(np.float64)
        n = normals[b, v].reshape(6, 3).astype(np.float64)
        scale = np.abs(p).max() * np.abs(n).max() * 100 + 1.0
        near = False
        for k in range(6):
            c = (corner2 if k < 3 else corner4)[b, v].astype(np.float64)
            s = 100.0 * float((p - c) @ n[k])
            near |= abs(s + 0.5) < 1e-5 * scale
        m = w2c[b, v].reshape(4, 4).astype(np.float64)
        cam = m @ np.append(p, 1.0)
        if abs(cam[2]) > 1e-9:
            u, w_ = cam[0] * fx / cam[2] + cx, cam[1] * fy / cam[2] + cy
            tol = 1e-5 * (abs(u) + abs(w_) + 50.0)
            near |= abs(u - np.floor(u) - 0.5) < tol or abs(w_ - np.floor(w_) - 0.5) < tol
            ui, vi = int(np.rint(u)), int(np.rint(w_))
            for du in (-1, 0, 1):
                for dv in (-1, 0, 1):
                    if 0 <= ui + du < W and 0 <= vi + dv < H:
                        z = float(depth[b, v, vi + dv, ui + du])
                        near |= abs(abs(z - cam[2]) - acc) < 1e-5 or abs(z - dmin) < 1e-6 or abs(z - dmax) < 1e-5
        ok[r] = near
    return ok
