"""Generates tests/golden/pc_util_r1.npz by importing the REFERENCE's own utils/pc_util.py (pure numpy, runs in the build
container) and calling point_cloud_label_to_surface_voxel_label_fast on seeded inputs.

    python tests/golden/make_golden_pc_util.py        (needs /root/reference; the .npz is committed)

Inputs are regenerated from the seeds by pc_util_inputs(); the fixture pins oracle/pc_util_ref.py (tests/test_golden.py,
CPU) and pn2_b200.pc_util (tests/test_voxel_gpu.py)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))

CASES = {"scannet_8192_r02": (11, 8192, 0.02), "scannet_4096_r0484": (12, 4096, 0.0484), "dup_2000_r05": (13, 2000, 0.05)}


def pc_util_inputs(name):
    """-> points (N, 6) float32 (xyz + rgb, as the evaluation loop passes them), label (N, 2) int64"""
    from pn2_b200 import scenes
    seed, n, _ = CASES[name]
    rng = np.random.default_rng(seed)
    pts = scenes.scannet_batch(500 + seed, 1, n)[0].astype(np.float32)
    if name.startswith("dup"):
        pts = pts[rng.integers(0, n // 5, n)]  # many exactly coincident points
    label = rng.integers(0, 21, (n, 2)).astype(np.int64)
    return pts, label


if __name__ == "__main__":
    import types
    for missing in ("plyfile", "matplotlib", "matplotlib.pyplot"):  # IO / plotting imports of the file, unused by the function
        try:
            __import__(missing)
        except ImportError:
            stub = types.ModuleType(missing)
            stub.PlyData = stub.PlyElement = None
            sys.modules[missing] = stub
    spec = importlib.util.spec_from_file_location("ref_pc_util", "/root/reference/utils/pc_util.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, (_, _, res) in CASES.items():
        pts, label = pc_util_inputs(name)
        uvidx, uvlabel, nvox = ref.point_cloud_label_to_surface_voxel_label_fast(pts, label, res=res)
        out[name + "/uvidx"], out[name + "/uvlabel"], out[name + "/nvox"] = uvidx, uvlabel, nvox
        print(name, uvidx.shape, uvidx.dtype, uvlabel.shape, nvox)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pc_util_r1.npz"), **out)
