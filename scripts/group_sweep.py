"""group_points over source-row lengths (N = 4 k .. 64 k): bit-equality with torch indexing and HBM fraction.
PN2_LIB_PATH selects a second build of the library for A/B runs.  python scripts/group_sweep.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np, torch
from pn2_b200 import _lib
if os.environ.get("PN2_LIB_PATH"):
    _lib.LIB_PATH = os.environ["PN2_LIB_PATH"]
from pn2_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(_pk))["hbm_gbs"] if os.path.exists(_pk) else 6650.0  # fallback: B200_PROFILING.md


def t_ms(fn, iters=7):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


g = torch.Generator(device=dev).manual_seed(0)
for (B, C, N, M, K) in [(32, 64, 8192, 1024, 32), (64, 64, 4096, 1024, 32), (64, 64, 16384, 4096, 32), (32, 64, 32768, 8192, 32),
                        (16, 64, 49152, 8192, 32), (16, 64, 65536, 16384, 32), (16, 128, 34720, 8192, 32), (3, 37, 20001, 1001, 7)]:
    f = torch.randn(B, C, N, device=dev)
    idx = torch.randint(0, N, (B, M, K), device=dev, dtype=torch.int32, generator=g)
    got = pu.grouping_operation(f, idx)
    want = torch.gather(f, 2, idx.view(B, 1, M * K).long().expand(B, C, M * K)).view(B, C, M, K)
    same = bool(torch.equal(got, want))
    del want
    ms = t_ms(lambda: pu.grouping_operation(f, idx))
    byts = B * (4 * M * K + 4 * C * min(N, M * K) + 4 * C * M * K)
    print(json.dumps({"op": "group_points", "B": B, "C": C, "N": N, "M": M, "K": K, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1),
                      "frac_hbm": round(byts / ms / 1e6 / PEAK, 3), "bit_equal": same}), flush=True)
    del f, idx, got
