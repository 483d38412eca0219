"""Generates tests/golden/ref_cuda_r1.npz on a B200: outputs of the REFERENCE's own CUDA kernels
(oracle/_ref/ref_cuda.so = utils/src/*_gpu.cu compiled verbatim, nvcc -O2, sm_100a) on seeded inputs.

    gpurun -- python tests/golden/make_golden.py        (writes gpurun_out/ref_cuda_r1.npz; copy it here)

Inputs are NOT stored: tests regenerate them from the same seeds through golden_inputs() below.
The fixtures pin the CPU oracle (tests/test_golden.py, no GPU needed) and the CUDA kernels
(tests/test_golden_gpu.py) to what the reference's implementation really computes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))

CASES = {
    # name: (kind, B, N, M)
    "scannet_2048": ("scannet", 2, 2048, 256),
    "dup_1024": ("dup", 2, 1024, 200),
    "lattice_600": ("lattice", 1, 600, 100),
    "uniform_8192": ("uniform", 1, 8192, 512),
    "dup_100": ("dup", 2, 100, 40),
    "uniform_3000": ("uniform", 1, 3000, 128),
}


def golden_inputs(name):
    from pn2_b200 import scenes
    kind, B, N, M = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    if kind == "scannet":
        xyz = np.stack([scenes.scannet_scene(7000 + b, N)[0] for b in range(B)])
    elif kind == "dup":
        xyz = np.empty((B, N, 3), np.float32)
        for b in range(B):
            pool = rng.random((max(N // 4, 1), 3)).astype(np.float32)
            xyz[b] = pool[rng.integers(0, pool.shape[0], N)]
    elif kind == "lattice":
        xyz = rng.integers(0, 6, (B, N, 3)).astype(np.float32) * np.float32(0.25)
    else:
        xyz = rng.random((B, N, 3)).astype(np.float32)
    feats = rng.standard_normal((B, 5, N)).astype(np.float32)
    radius = {"scannet": 0.2, "dup": 0.15, "lattice": 0.3, "uniform": 0.1}[kind]
    return xyz, feats, M, radius


def main():
    import torch
    from oracle import ref_cuda
    assert ref_cuda.available(), "needs oracle/_ref/ref_cuda.so and a GPU"
    dev = torch.device("cuda:0")
    out = {}
    for name in CASES:
        xyz, feats, M, radius = golden_inputs(name)
        x = torch.from_numpy(xyz).to(dev)
        f = torch.from_numpy(feats).to(dev)
        fps = ref_cuda.furthest_point_sample(x, M)
        new_xyz = ref_cuda.gather_operation(x.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        bq = ref_cuda.ball_query(radius, 16, x, new_xyz)
        grouped = ref_cuda.grouping_operation(f, bq)
        dist, nn_idx = ref_cuda.three_nn(x, new_xyz)
        d = dist.clone()
        d[d < 1e-10] = 1e-10
        w = 1.0 / d
        w = w / w.sum(-1, keepdim=True)
        interp = ref_cuda.three_interpolate(f[:, :, :M].contiguous(), nn_idx, w)
        out[name + "/fps"] = fps.cpu().numpy()
        out[name + "/ball"] = bq.cpu().numpy()
        out[name + "/grouped_sum"] = grouped.double().sum((2, 3)).cpu().numpy()  # checksum: the gather itself is exact
        out[name + "/nn_idx"] = nn_idx.cpu().numpy()
        out[name + "/nn_dist"] = dist.cpu().numpy()
        out[name + "/weight"] = w.cpu().numpy()
        out[name + "/interp"] = interp.cpu().numpy().astype(np.float32)
    dst = os.path.join(ROOT, "gpurun_out", "ref_cuda_r1.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
