"""Generates tests/golden/rpn_r1.npz by importing the REFERENCE's own model/pointmaskrcnn.py (its `pointnet2_cuda` import is
satisfied by this repo's drop-in module; iou_spheres and nms are pure torch and run on the CPU of the build container).

    python tests/golden/make_golden_rpn.py        (needs /root/reference; the .npz is committed)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))

CASES = {"mixed_200": (21, 200, 0.3), "dense_500": (22, 500, 0.7), "nested_64": (23, 64, 0.1), "sparse_300": (24, 300, 0.5)}


def rpn_inputs(name):
    """-> spheres (N, 4) float32, scores (N,) float32 (distinct values: torch.sort leaves ties unspecified)"""
    seed, n, _ = CASES[name]
    rng = np.random.default_rng(seed)
    spread = {"mixed": 6.0, "dense": 3.0, "nested": 0.5, "sparse": 40.0}[name.split("_")[0]]
    c = rng.uniform(-spread, spread, (n, 3))
    r = rng.uniform(0.3, 2.5, (n, 1))
    if name.startswith("nested"):
        c[n // 2:] = c[:n - n // 2] + rng.normal(0, 0.02, (n // 2, 3))  # concentric pairs: the "inside" branch
    spheres = np.concatenate([c, r], 1).astype(np.float32)
    scores = rng.permutation(n).astype(np.float32) / n
    return spheres, scores


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    import importlib
    ref = importlib.import_module("model.pointmaskrcnn")
    out = {}
    for name, (_, _, thr) in CASES.items():
        s, sc = rpn_inputs(name)
        st, sct = torch.from_numpy(s), torch.from_numpy(sc)
        iou = ref.iou_spheres(st, st, no_grad=True).numpy()
        keep = ref.nms(st, sct, threshold=thr).numpy()
        out[name + "/iou"], out[name + "/keep"] = iou, keep
        print(name, iou.shape, "kept", len(keep), "nonzero iou", float((iou > 0).mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rpn_r1.npz"), **out)
