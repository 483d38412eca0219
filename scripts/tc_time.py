"""Per-launch timing of the eight tensor-core MLP launches of one SSG forward (developer timing).

Single-stream eager forwards in bf16; every pn2_*_mlp*_bf16 call is bracketed by CUDA events; prints the median over the
timed forwards per launch and the sum, plus max|diff| against the fp32 fused path (2e-2 bar of the bf16 path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pn2_b200 import _lib, scenes  # noqa: E402
import pn2_b200.pointnet_util as U  # noqa: E402
from pn2_b200.models import PointNet2SemSeg  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32"))
ITERS = int(os.environ.get("ITERS", "30" if os.environ.get("PRECISION", "bf16") == "bf16" else "8"))
torch.manual_seed(0)
model = PointNet2SemSeg(21).eval().to(dev)
model.single_stream = True
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
x6 = pts.permute(0, 2, 1).contiguous()
if os.environ.get("PN2_LIB_PATH"):  # A/B runs: a second build of the library (e.g. the previous commit's)
    _lib.LIB_PATH = os.environ["PN2_LIB_PATH"]
lib = _lib.load()
if os.environ.get("TC_MAX_CTAS"):
    lib.pn2_debug_set_tc_max_ctas(int(os.environ["TC_MAX_CTAS"]))
if os.environ.get("TC_WORKERS"):
    lib.pn2_debug_set_tc_workers(int(os.environ["TC_WORKERS"]))
PREC = os.environ.get("PRECISION", "bf16")  # fp32: times the FFMA kernels of row_mlp.cu instead
TC = ("pn2_sa_mlp_max_bf16", "pn2_fp_mlp_bf16") if PREC == "bf16" else ("pn2_sa_mlp_max", "pn2_fp_mlp")
NAMES = ["sa1", "sa2", "sa3", "sa4", "fp4", "fp3", "fp2", "fp1+head"]
orig_call = _lib.call
rec = []


def hooked(name, *args):
    if name in TC and rec is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig_call(name, *args)
        b.record()
        rec.append((a, b))
        return r
    return orig_call(name, *args)


_lib.call = hooked
U._lib.call = hooked
with torch.no_grad():
    U.set_mlp_precision("fp32")
    want = model(x6[:, :3], x6[:, 3:]).float()
    U.set_mlp_precision(PREC)
    for _ in range(25 if PREC == "bf16" else 5):  # clocks and caches settle (the first process on a fresh box reads ~10 % slow otherwise)
        got = model(x6[:, :3], x6[:, 3:])
    torch.cuda.synchronize()
    rec.clear()
    for _ in range(ITERS):
        model(x6[:, :3], x6[:, 3:])
    torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in rec]).reshape(ITERS, -1)
med = np.median(ms, axis=0) * 1e3
for nm, v in zip(NAMES, med):
    print("%-9s %7.1f us" % (nm, v))
print("sum       %7.1f us" % med.sum())
err = (got.float() - want).abs().max().item() / want.abs().max().item()
print("bf16 vs fp32 fused path: max|diff| / max|ref| = %.3e" % err)
