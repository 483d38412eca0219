"""Shared by the CPU and GPU module parity tests: the committed vectors produced by the REFERENCE's own module classes
(tests/golden/modules_r2.npz, made by tests/golden/make_golden_modules.py) and the seeded problems they belong to."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODS = os.path.join(HERE, "golden", "modules_r2.npz")
if os.path.join(HERE, "golden") not in sys.path:
    sys.path.insert(0, os.path.join(HERE, "golden"))
_spec = importlib.util.spec_from_file_location("make_golden_modules", os.path.join(HERE, "golden", "make_golden_modules.py"))
mgm = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mgm)


def unit_seed(name):
    return 100 + sorted(mgm.UNIT_CASES).index(name)


def check_sample(got, gold, key, rel, axis=1, stride=None):
    """got: full output (numpy); compares the strided sample, the global fp64 sum and the shape stored for `key`."""
    stride = mgm.STRIDE if stride is None else stride
    assert tuple(gold[key + "/shape"]) == got.shape
    want = gold[key + "/sample"]
    sl = [slice(None)] * got.ndim
    sl[axis] = slice(None, None, stride)
    scale = float(gold[key + "/absmax"])
    err = float(np.abs(got[tuple(sl)] - want).max())
    assert err <= rel * scale, "%s: max abs err %.3e vs %.1e * %.3e" % (key, err, rel, scale)
    assert abs(float(got.astype(np.float64).sum()) - float(gold[key + "/sum"])) <= rel * scale * got.size * 0.05 + 1e-3
    return err / scale
