"""Generates tests/golden/projection_r2.npz by running the REFERENCE's own utils/projection.py (ProjectionHelper :5-230,
Projection :234-267) and the view reductions of model/pointnet2multiview.py (:30-43 max pool, :83-102 first non-zero view)
on the CPU of the build container (tests/golden/ref_harness.py: `.cuda()` resolves to the host, nothing else changes).

    python tests/golden/make_golden_projection.py        (needs /root/reference; the .npz is committed)

The reference's arithmetic runs on whatever BLAS backs torch.mm / torch.bmm / torch.inverse; here that is the host's.  The
reference on a GPU (cuBLAS) differs from these vectors in the last ulp of a dot product, i.e. in the decision for a point
that sits on a rounding boundary -- tests name those cases explicitly instead of allowing a blanket tolerance.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))

from pn2_b200 import scenes  # noqa: E402  (seeded numpy generators only; no product kernels involved)

# the literals of train_scannet_multiview_semseg.py:109-110 -- a Python list of doubles, as the reference passes it
INTRINSIC = [[37.01983, 0, 20, 0], [0, 38.52470, 15.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
DEPTH_MIN, DEPTH_MAX, IMAGE_DIMS, ACCURACY = 0.1, 4.0, [41, 32], 0.05

# name: (first scene id, B, N, V, C)
CASES = {"v3_n8192": (40, 2, 8192, 3, 128), "v5_n8192": (42, 1, 8192, 5, 128), "v5_n3000": (43, 1, 3000, 5, 16)}
KEEP_POINTS = 64          # full feature columns kept for the first points of every cloud (the rest: sha256 + channel sums)
FRUSTUM_CASE = (90, 8192, 150, 5)   # scene id, N, poses, rng seed (row N2: the loader's best-view selection)


def lifting_inputs(name):
    """-> xyz (B,N,3), feats (B,V,C,H,W), depth (B,V,H,W), poses (B,V,4,4); view 0 carries an all-zero image row so the
    first-non-zero reduction has visible-but-zero columns to replace (model/pointnet2multiview.py:96-98)."""
    first, B, N, V, C = CASES[name]
    xyz, feats, depth, poses = [], [], [], []
    for b in range(B):
        x, _ = scenes.scannet_scene(first + b, N)
        f, d, p = scenes.multiview_inputs(first + b, x, V, C)
        f[0, :, 5, :] = 0.0
        f[0, :, 17, 10:30] = 0.0
        xyz.append(x); feats.append(f); depth.append(d); poses.append(p)
    return np.stack(xyz), np.stack(feats), np.stack(depth), np.stack(poses)


def frustum_inputs():
    """-> points (N,3) float32, poses (P,4,4) float32: cameras scattered around one scene, many of them seeing little."""
    sid, N, P, seed = FRUSTUM_CASE
    x, _ = scenes.scannet_scene(sid, N)
    rng = np.random.default_rng(seed)
    centre = x.mean(0)
    poses = []
    for q, (a, d, h) in enumerate(zip(rng.uniform(0, 6.28, P), rng.uniform(0.5, 4.0, P), rng.uniform(0.0, 2.0, P))):
        eye = centre + np.array([np.cos(a) * d, np.sin(a) * d, h])
        target = centre + rng.normal(0, 0.5, 3)
        if q % 15 == 7:
            target = eye + (eye - centre) + rng.normal(0, 0.3, 3)   # looking away: sees (almost) nothing
        poses.append(scenes.look_at_pose(eye, target))
    return x, np.stack(poses)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reduce_max(maps, num_images):
    """model/pointnet2multiview.py:36-41, the statements of PointNet2Multiview.forward on the stacked per-view maps."""
    imageft = torch.stack(maps, dim=2)
    sz = imageft.shape
    imageft = imageft.view(sz[0], -1, num_images)
    imageft = torch.nn.functional.max_pool1d(imageft, kernel_size=num_images)
    return imageft.view(sz[0], sz[1])


def reduce_first(maps, num_images):
    """model/pointnet2multiview.py:89-99 (PointNet2Multiview2.forward); the literal 128 of :97 is the channel count there."""
    imageft = torch.stack(maps, dim=2)
    sz = imageft.shape
    imageft = imageft.view(sz[0], -1, num_images)
    for j in range(imageft.shape[2]):
        if j == 0:
            imageft_final = imageft[:, :, j]
        else:
            mask = ((imageft_final == 0).sum(0) == sz[0]).nonzero().squeeze(1)
            imageft_final[:, mask] = imageft[:, mask, j]
    return imageft_final.reshape(sz[0], sz[1])


if __name__ == "__main__":
    from ref_harness import reference_on_cpu
    out = {}
    with reference_on_cpu():
        from utils.projection import Projection, ProjectionHelper
        helper = ProjectionHelper(INTRINSIC, DEPTH_MIN, DEPTH_MAX, IMAGE_DIMS, ACCURACY)
        for name, (_, B, N, V, C) in CASES.items():
            xyz, feats, depth, poses = lifting_inputs(name)
            corners = np.zeros((B, V, 8, 4), np.float32)
            normals = np.zeros((B, V, 6, 3), np.float32)
            w2c = np.zeros((B, V, 4, 4), np.float32)
            pix = np.full((B, V, N), -1, np.int32)
            frustum = np.zeros((B, V, N), bool)
            for b in range(B):
                pts = torch.from_numpy(xyz[b])
                maps = []
                for v in range(V):
                    c2w = torch.from_numpy(poses[b, v])
                    cc = helper.compute_frustum_corners(c2w)
                    nr = helper.compute_frustum_normals(cc)
                    corners[b, v] = cc[:, :, 0].numpy()
                    normals[b, v] = nr.numpy()
                    w2c[b, v] = torch.inverse(c2w).numpy()          # utils/projection.py:178
                    frustum[b, v] = helper.points_in_frustum(cc, nr, pts, return_mask=True).numpy()
                    res = helper.compute_projection(pts, torch.from_numpy(depth[b, v]), c2w, N)
                    assert res is not None, "the synthetic views must see the scene"
                    ind3d, ind2d = res
                    n = int(ind3d[0])
                    assert n == int(ind2d[0]) and n > 0
                    pix[b, v, ind3d[1:1 + n].numpy()] = ind2d[1:1 + n].numpy()
                    maps.append(Projection.apply(torch.from_numpy(feats[b, v]), ind3d, ind2d, N))
                for red, fn in (("max", reduce_max), ("first", reduce_first)):
                    o = fn([m.clone() for m in maps], V).numpy()
                    assert o.shape == (C, N)
                    out["%s/%s/sha/%d" % (name, red, b)] = np.array(sha(o))
                    out["%s/%s/chansum/%d" % (name, red, b)] = o.astype(np.float64).sum(1)
                    out["%s/%s/head/%d" % (name, red, b)] = o[:, :KEEP_POINTS].copy()
                # a single view through Projection.forward alone (row A13)
                out["%s/single/sha/%d" % (name, b)] = np.array(sha(maps[V - 1].numpy()))
            out[name + "/corners"], out[name + "/normals"], out[name + "/w2c"] = corners, normals, w2c
            out[name + "/pix"] = pix
            out[name + "/frustum"] = np.packbits(frustum, axis=-1)
            print(name, "lifted fraction per view", (pix >= 0).mean(-1).round(3).tolist(),
                  "in frustum", frustum.mean(-1).round(3).tolist())

        # row N2: the loader's frustum count, fp64 on the host, data_utils/ScanNetDataLoader.py:91-97
        x, poses = frustum_inputs()
        P = poses.shape[0]
        masks = np.zeros((P, x.shape[0]), bool)
        counts = np.zeros(P, np.int64)
        fcorners, fnormals = np.zeros((P, 8, 3), np.float32), np.zeros((P, 6, 3), np.float32)
        with torch.no_grad():
            for q in range(P):
                cc = helper.compute_frustum_corners(torch.from_numpy(poses[q]))[:, :3, 0]
                nr = helper.compute_frustum_normals(cc)
                fcorners[q], fnormals[q] = cc.numpy(), nr.numpy()
                pts64 = torch.DoubleTensor(x.astype(np.float64))
                counts[q] = int(helper.points_in_frustum_cpu(cc.double(), nr.double(), pts64))
                masks[q] = helper.points_in_frustum_cpu(cc.double(), nr.double(), pts64, return_mask=True).numpy()
        # the loader's selection rule, :98-105, for num_images = 5
        poseDict = {q: int(counts[q]) for q in range(P)}
        frame_ids = []
        for i in range(5):
            maximum = max(poseDict, key=poseDict.get)
            if (i == 0) | (poseDict[maximum] > 100):
                frame_ids.append(int(maximum))
                del poseDict[maximum]
            else:
                frame_ids.append(frame_ids[0])
        out["frustum/corners"], out["frustum/normals"] = fcorners, fnormals
        out["frustum/counts"], out["frustum/masks"], out["frustum/best5"] = counts, np.packbits(masks, axis=-1), np.array(frame_ids)
        print("frustum counts: max", counts.max(), "zeros", int((counts == 0).sum()), "best5", frame_ids)
    np.savez_compressed(os.path.join(HERE, "projection_r2.npz"), **out)
    print("wrote projection_r2.npz", os.path.getsize(os.path.join(HERE, "projection_r2.npz")), "bytes")
