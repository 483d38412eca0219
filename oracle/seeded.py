"""Seeded parameter values for parity tests (TEST INFRASTRUCTURE ONLY).

The golden generators fill the REFERENCE's module classes with these values and the tests fill this repo's drop-in
classes (same constructor arguments, same state_dict keys -- model/pointnet_util.py:70-83,114-131,174-183) with the same
ones, so a fixture does not depend on torch's RNG stream or on parameter creation order.  BatchNorm running statistics
and affine parameters are non-trivial on purpose: eval-mode folding must be exercised.
"""
import numpy as np
import torch


def fill_seeded(module, seed):
    """Overwrites every entry of module.state_dict() in sorted-key order from numpy's PCG64(seed).  Returns module."""
    rng = np.random.default_rng(int(seed))
    sd = module.state_dict()
    new = {}
    for key in sorted(sd):
        t = sd[key]
        shape = tuple(t.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            new[key] = torch.zeros_like(t)
            continue
        if leaf == "running_var":
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf == "running_mean":
            a = rng.normal(0.0, 0.1, shape)
        elif leaf == "weight" and t.dim() == 1:            # BatchNorm gamma
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf == "weight":                              # conv (cout, cin, 1[, 1])
            a = rng.normal(0.0, 1.0, shape) * np.sqrt(2.0 / max(shape[1], 1))
        else:                                               # conv bias / BatchNorm beta
            a = rng.normal(0.0, 0.1, shape)
        new[key] = torch.from_numpy(np.asarray(a, dtype=np.float32)).reshape(shape)
    module.load_state_dict(new)
    return module
