"""Phase timeline of every tensor-core MLP launch of one SSG forward (developer profiling).

Arms pn2_debug_set_tc_timestamps for the k-th tcgen05 launch of a single-stream eager forward and prints, for CTA 0:
worker stamps (tile start, gather done, then per layer accumulators-ready / epilogue-done) and the MMA issuer's stamps
(operand seen, MMAs issued + committed).  All clock64 of the same SM.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch  # noqa: E402
from pn2_b200 import _lib, scenes  # noqa: E402
from pn2_b200.models import PointNet2SemSeg  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32"))
torch.manual_seed(0)
model = PointNet2SemSeg(21).eval().to(dev)
model.single_stream = True
pts = torch.from_numpy(scenes.scannet_batch(0, B, 8192)).to(dev)
x6 = pts.permute(0, 2, 1).contiguous()
lib = _lib.load()
lib.pn2_debug_set_tc_timestamps.argtypes = [ctypes.c_void_p]
if os.environ.get("TC_MAX_CTAS"):
    lib.pn2_debug_set_tc_max_ctas(int(os.environ["TC_MAX_CTAS"]))
TC = ("pn2_sa_mlp_max_bf16", "pn2_fp_mlp_bf16")
NAMES = ["sa1", "sa2", "sa3", "sa4", "fp4", "fp3", "fp2", "fp1+head"]

orig_call = _lib.call
state = {"k": -1, "seen": 0, "buf": None}


def hooked(name, *args):
    if name in TC:
        if state["seen"] == state["k"]:
            lib.pn2_debug_set_tc_timestamps(state["buf"].data_ptr())
        state["seen"] += 1
    return orig_call(name, *args)


_lib.call = hooked
import pn2_b200.pointnet_util as U  # noqa: E402
U._lib.call = hooked
U.set_mlp_precision("bf16")  # the drop-in modules default to the fp32 path

with torch.no_grad():
    for _ in range(2):
        model(x6[:, :3], x6[:, 3:])
    for k, nm in enumerate(NAMES):
        buf = torch.zeros(512, dtype=torch.int64, device=dev)
        state.update(k=k, seen=0, buf=buf)
        torch.cuda.synchronize()
        model(x6[:, :3], x6[:, 3:])
        torch.cuda.synchronize()
        t = buf.cpu().tolist()
        w = [v for v in t[:256] if v]
        m = [v for v in t[256:384] if v]
        ww = [v - 1 for v in t[384:] if v]
        if not w:
            print(nm, "no stamps")
            continue
        t0 = w[0]
        print("== %s: worker stamps %d, issuer stamps %d" % (nm, len(w), len(m)))
        print("  worker (rel):", [v - t0 for v in w[:30]])
        print("  issuer (rel):", [v - t0 for v in m[:30]])
        print("  issuer issue-phase lengths:", [b - a for a, b in zip(m[0:30:2], m[1:30:2])])
        print("  of which waiting for weights:", ww[:15])
