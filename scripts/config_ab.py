"""The `extra` legs of bench.py (configs 2-4) against a chosen build of the library: PN2_LIB_PATH=... python scripts/config_ab.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import _lib
if os.environ.get("PN2_LIB_PATH"):
    _lib.LIB_PATH = os.environ["PN2_LIB_PATH"]
import bench
from pn2_b200 import pointnet_util
pointnet_util.set_mlp_precision(os.environ.get("PRECISION", "bf16"))  # what bench.py measures by default
out = bench.time_other_configs(torch.device("cuda:0"), 32)
print(json.dumps({k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "what"} for k, v in out.items()}))
