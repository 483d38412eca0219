"""Generates tests/golden/modules_r2.npz by running the REFERENCE's own, unmodified module classes
(model/pointnet_util.py:70-221 PointNetSetAbstraction / PointNetSetAbstractionMsg / PointNetFeaturePropagation,
model/pointnet2.py:131-162 PointNet2SemSeg, model/pointmaskrcnn.py:8-32 PointNet2, model/pointnet2multiview.py
PointNet2Multiview2 :61-121 and PointNet2Multiview2Msg :179-233, utils/projection.py) on the CPU of the build container
through tests/golden/ref_harness.py: `pointnet2_cuda` is backed by the C oracle (itself pinned bit-exact to the reference's
compiled kernels), `.cuda()` resolves to the host, ENet is the identity (its output feature maps are the inputs here).

    python tests/golden/make_golden_modules.py        (needs /root/reference; the .npz is committed; ~2 min)

Parameters come from oracle/seeded.py (numpy PCG64 by state_dict key), inputs from pn2_b200/scenes.py (seeded numpy), so
the tests rebuild exactly the same problem without the reference.  Large outputs are stored as a strided sample plus fp64
channel sums.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))

from oracle.seeded import fill_seeded  # noqa: E402
from pn2_b200 import scenes  # noqa: E402

NUM_CLASSES = 21

# unit cases: name -> (kind, constructor args, B, N, D (feature channels), extra)
UNIT_CASES = {
    "sa_ssg": ("sa", (256, 0.2, 32, 6 + 3, [32, 32, 64], False), 2, 2048, 6, None),
    "sa_ssg_noD": ("sa", (128, 0.3, 16, 0 + 3, [16, 32], False), 2, 1000, 0, None),
    "sa_group_all": ("sa", (None, None, None, 5 + 3, [32, 64], True), 2, 512, 5, None),
    "sa_msg": ("msg", (128, [0.1, 0.2], [16, 32], 6, [[16, 16, 32], [32, 32, 64]]), 2, 2048, 6, None),
    "sa_msg_noD": ("msg", (64, [0.2, 0.4], [8, 32], 0, [[16, 32], [16, 48]]), 2, 1024, 0, None),
    "fp_skip": ("fp", (16 + 64, [64, 32]), 2, 2048, 16, (256, 64)),      # extra = (S, D2)
    "fp_noskip": ("fp", (128, [128, 64]), 2, 500, 0, (100, 128)),
    "fp_s1": ("fp", (4 + 16, [32]), 2, 300, 4, (1, 16)),
}
STRIDE = 8   # whole-network outputs keep every STRIDE-th point


def scene_inputs(first_scene, B, N, D, seed):
    """-> xyz (B,3,N) float32 (ScanNet-shaped crops, drawn with replacement: duplicate points, FPS ties), feat (B,D,N) or None"""
    pts = scenes.scannet_batch(first_scene, B, N)
    xyz = np.ascontiguousarray(pts[:, :, :3].transpose(0, 2, 1))
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((B, D, N)).astype(np.float32) if D else None
    return xyz, feat


def unit_inputs(name):
    kind, args, B, N, D, extra = UNIT_CASES[name]
    seed = 500 + sorted(UNIT_CASES).index(name)
    xyz, feat = scene_inputs(600 + 10 * sorted(UNIT_CASES).index(name), B, N, D, seed)
    if kind != "fp":
        return (xyz, feat)
    S, D2 = extra
    rng = np.random.default_rng(seed + 1000)
    xyz2 = np.ascontiguousarray(xyz[:, :, rng.permutation(N)[:S]])
    p2 = rng.standard_normal((B, D2, S)).astype(np.float32)
    return (xyz, xyz2, feat, p2)


def semseg_inputs(B=2):
    pts = scenes.scannet_batch(700, B, 8192)
    x = np.ascontiguousarray(pts.transpose(0, 2, 1))
    return x[:, :3].copy(), x[:, 3:].copy()


def backbone_inputs(N):
    xyz, feat = scenes.lidar_sweep(710, N)
    return np.ascontiguousarray(xyz.T[None]), np.ascontiguousarray(feat.T[None])


def multiview_inputs(B, V):
    """-> xyz (B,3,N), feats (B,V,128,32,41), depth (B,V,32,41), poses (B,V,4,4) as the multi-view train loop holds them"""
    N = 8192
    xs, fs, ds, ps = [], [], [], []
    for b in range(B):
        x, _ = scenes.scannet_scene(720 + b, N)
        f, d, p = scenes.multiview_inputs(720 + b, x, V, 128)
        f[0, :, 9, :] = 0.0
        xs.append(x); fs.append(f); ds.append(d); ps.append(p)
    return np.ascontiguousarray(np.stack(xs).transpose(0, 2, 1)), np.stack(fs), np.stack(ds), np.stack(ps)


def T(a):
    return None if a is None else torch.from_numpy(a)


def store_big(out, key, y):
    """y (B, N, C) or (B, C, N) numpy -> strided sample along the point axis + fp64 sums"""
    out[key + "/shape"] = np.array(y.shape)
    out[key + "/sum"] = y.astype(np.float64).sum()
    out[key + "/absmax"] = np.abs(y).max()
    return y


if __name__ == "__main__":
    from ref_harness import reference_on_cpu
    import make_golden_projection as mgj
    out = {}
    torch.set_num_threads(os.cpu_count() or 1)
    with reference_on_cpu(), torch.no_grad():
        from model.pointnet_util import PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg
        from model.pointnet2 import PointNet2SemSeg
        from model.pointmaskrcnn import PointNet2
        from model.pointnet2multiview import PointNet2Multiview2, PointNet2Multiview2Msg
        from utils.projection import ProjectionHelper
        ctor = {"sa": PointNetSetAbstraction, "msg": PointNetSetAbstractionMsg, "fp": PointNetFeaturePropagation}
        for i, name in enumerate(sorted(UNIT_CASES)):
            kind, args, B, N, D, extra = UNIT_CASES[name]
            mod = fill_seeded(ctor[kind](*args), 100 + i).eval()
            inp = [T(a) for a in unit_inputs(name)]
            res = mod(*inp)
            if kind == "fp":
                out[name + "/out"] = res.numpy()
            else:
                out[name + "/new_xyz"], out[name + "/out"] = res[0].numpy(), res[1].numpy()
            print(name, [tuple(r.shape) for r in (res if isinstance(res, tuple) else (res,))])
            if name == "sa_ssg":   # the same block with batch statistics (training-mode BatchNorm2d, model/pointnet_util.py:107)
                mod.train()
                res = mod(*inp)
                out[name + "/train_out"] = res[1].numpy()
                out[name + "/train_running_mean0"] = mod.mlp_bns[0].running_mean.numpy().copy()
                out[name + "/train_running_var0"] = mod.mlp_bns[0].running_var.numpy().copy()

        # ---- config 1: PointNet2SemSeg at B=2, N=8192 (model/pointnet2.py:131-162), eval and train-mode BatchNorm ----
        xyz, rgb = semseg_inputs()
        net = fill_seeded(PointNet2SemSeg(NUM_CLASSES), 200).eval()
        y = net(T(xyz), T(rgb)).numpy()
        store_big(out, "semseg_eval", y)
        out["semseg_eval/sample"] = y[:, ::STRIDE].copy()
        out["semseg_eval/argmax"] = y.argmax(-1).astype(np.uint8)
        print("semseg eval", y.shape, float(np.abs(y).max()))
        net.train()
        net.drop1.eval()   # dropout is the only random op of the forward; batch statistics are what this case pins
        y = net(T(xyz), T(rgb)).numpy()
        store_big(out, "semseg_train", y)
        out["semseg_train/sample"] = y[:, ::STRIDE].copy()
        out["semseg_train/sa1_bn0_running_mean"] = net.sa1.mlp_bns[0].running_mean.numpy().copy()
        out["semseg_train/fp1_bn2_running_var"] = net.fp1.mlp_bns[2].running_var.numpy().copy()
        print("semseg train", y.shape)

        # ---- config 4: the nuScenes backbone (model/pointmaskrcnn.py:8-32) on a 16384-point synthetic sweep ----
        bx, bf = backbone_inputs(16384)
        bb = fill_seeded(PointNet2(), 210).eval()
        y = bb(T(bx), T(bf)).numpy()       # (1, 128, N)
        store_big(out, "backbone16k", y)
        out["backbone16k/sample"] = y[:, :, ::16].copy()
        print("backbone", y.shape)

        # ---- config 3: PointNet2Multiview2, lifting + point branch, as train_scannet_multiview_semseg.py:138-163 drives it ----
        helper = ProjectionHelper(mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS, mgj.ACCURACY)
        for cls, tag, B, V, seed in ((PointNet2Multiview2, "mv2_first", 2, 3, 220), (PointNet2Multiview2Msg, "mv2msg_max", 1, 5, 230)):
            mx, mf, md, mp = multiview_inputs(B, V)
            N = mx.shape[2]
            ind3d, ind2d = [], []
            for b in range(B):
                pts = T(np.ascontiguousarray(mx[b].T))
                pairs = [helper.compute_projection(pts, T(md[b, v]), T(mp[b, v]), N) for v in range(V)]
                assert None not in pairs
                a3, a2 = zip(*pairs)
                ind3d.append(torch.stack(a3)); ind2d.append(torch.stack(a2))
            net = fill_seeded(cls(NUM_CLASSES), seed).eval()
            captured = {}
            net.sa1_feat.register_forward_pre_hook(lambda m, args: captured.__setitem__("img", args[1].detach().clone()))
            y = net(T(mx), [T(mf[b]) for b in range(B)], ind3d, ind2d).numpy()
            img = captured["img"].numpy()    # (B, 128, N): the lifted, view-reduced image features
            store_big(out, tag, y)
            out[tag + "/sample"] = y[:, ::STRIDE].copy()
            out[tag + "/image_features_sha"] = np.array([mgj.sha(img[b]) for b in range(B)])
            out[tag + "/image_features_sum"] = img.astype(np.float64).sum((1, 2))
            print(tag, y.shape, "lifted nonzero fraction", float((img != 0).any(1).mean()))
    path = os.path.join(HERE, "modules_r2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
