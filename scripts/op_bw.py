"""HBM-roofline check of the standalone gather / interpolate operators at the C1 and sweep shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def t_ms(fn, iters=7):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
g = torch.Generator(device=dev).manual_seed(0)
for (B, C, m, n) in [(32, 128, 1024, 8192), (64, 128, 2048, 8192), (64, 128, 4096, 16384), (16, 256, 256, 1024)]:
    f = torch.randn(B, C, m, device=dev); idx = torch.randint(0, m, (B, n, 3), device=dev, dtype=torch.int32, generator=g)
    w = torch.rand(B, n, 3, device=dev); w = (w / w.sum(-1, keepdim=True)).contiguous()
    ms = t_ms(lambda: pu.three_interpolate(f, idx, w)); byts = B * (24 * n + 4 * C * m + 4 * C * n)
    print("three_interpolate B=%d C=%d m=%d n=%d: %.3f ms  %.0f GB/s  %.1f%% of HBM" % (B, C, m, n, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / PEAK))
for (B, C, N, M, K) in [(32, 64, 8192, 1024, 32), (32, 3, 8192, 1024, 32), (64, 128, 16384, 4096, 32), (32, 128, 1024, 256, 32)]:
    f = torch.randn(B, C, N, device=dev); idx = torch.randint(0, N, (B, M, K), device=dev, dtype=torch.int32, generator=g)
    ms = t_ms(lambda: pu.grouping_operation(f, idx)); byts = B * (4 * M * K + 4 * C * min(N, M * K) + 4 * C * M * K)
    print("group_points B=%d C=%d N=%d M=%d K=%d: %.3f ms  %.0f GB/s  %.1f%% of HBM" % (B, C, N, M, K, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / PEAK))
for (B, C, N, M) in [(64, 128, 16384, 4096), (32, 3, 8192, 1024)]:
    f = torch.randn(B, C, N, device=dev); idx = torch.randint(0, N, (B, M), device=dev, dtype=torch.int32, generator=g)
    ms = t_ms(lambda: pu.gather_operation(f, idx)); byts = B * (4 * M + 8 * C * M)
    print("gather_points B=%d C=%d N=%d M=%d: %.3f ms  %.0f GB/s  %.1f%% of HBM" % (B, C, N, M, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / PEAK))
