"""Graphed / eager MSG train step (bench.py's config-2 leg) against a chosen build of the library:
PN2_LIB_PATH=... python scripts/train_ab.py [scenes per step]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
import torch
from pn2_b200 import _lib
if os.environ.get("PN2_LIB_PATH"):
    _lib.LIB_PATH = os.environ["PN2_LIB_PATH"]
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g, e = bench.time_train_step_msg(torch.device("cuda:0"), batch=B)
print("MSG train step, %d scenes: graphed %.2f ms, eager %.2f ms" % (B, g, e))
