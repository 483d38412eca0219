"""BASELINE config 5: geometry-op microbench sweep (FPS / ball query / three_nn / three_interpolate / group), batch 64,
uniform clouds scaled so that a ball of radius r holds ~16 points.  CUDA events, 3 warm-ups, L2 flushed between
iterations.  Prints one JSON object per (op, N)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import pointnet2_utils as pu, scenes
from pn2_b200.pointnet_util import fps_gather_cl

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
B = 64


def t_ms(fn, iters=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


sizes = [int(s) for s in sys.argv[1:]] or [4096, 8192, 16384, 32768]
for N in sizes:
    seed = 2000 + int(np.log2(N))
    xyz = torch.from_numpy(np.stack([scenes.uniform_cloud(seed * 100 + b, N) for b in range(B)])).to(dev)
    M = N // 4
    r = float((16.0 / N / (4.0 / 3.0 * np.pi)) ** (1.0 / 3.0))  # ~16 neighbours in the unit cube
    ms = t_ms(lambda: fps_gather_cl(xyz, M), iters=3)
    print(json.dumps({"op": "fps", "N": N, "M": M, "B": B, "ms": ms, "us_per_round": ms * 1e3 / (M - 1),
                      "updates_per_s": B * float(N) * (M - 1) / ms * 1e3}))
    _, new_xyz = fps_gather_cl(xyz, M)
    ms = t_ms(lambda: pu.ball_query(r, 32, xyz, new_xyz))
    print(json.dumps({"op": "ball_query", "N": N, "M": M, "K": 32, "radius": r, "ms": ms,
                      "queries_per_s": B * M / ms * 1e3, "gbs": B * (12 * N + 12 * M + 4 * M * 32) / ms / 1e6}))
    ms = t_ms(lambda: pu.three_nn(xyz, new_xyz))
    print(json.dumps({"op": "three_nn", "n": N, "m": M, "ms": ms, "points_per_s": B * N / ms * 1e3}))
    d, idx = pu.three_nn(xyz, new_xyz)
    w = torch.full_like(d, 1.0 / 3.0)
    feats = torch.randn(B, 128, M, device=dev)
    ms = t_ms(lambda: pu.three_interpolate(feats, idx, w))
    byts = B * (24 * N + 4 * 128 * M + 4 * 128 * N)
    print(json.dumps({"op": "three_interpolate", "C": 128, "m": M, "n": N, "ms": ms, "gbs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / PEAK}))
    bq = pu.ball_query(r, 32, xyz, new_xyz)
    f = torch.randn(B, 64, N, device=dev)
    ms = t_ms(lambda: pu.grouping_operation(f, bq))
    byts = B * (4 * M * 32 + 4 * 64 * N + 4 * 64 * M * 32)
    print(json.dumps({"op": "group_points", "C": 64, "N": N, "M": M, "K": 32, "ms": ms, "gbs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / PEAK}))
    del xyz, feats, f, bq, idx, w, d
    torch.cuda.empty_cache()
