"""N > 1 path on the CPU (gloo, world_size 2): scenes are sharded across ranks with NO data-path collective
(SURVEY.md 8e); the only exchange is the max-over-ranks of the timing, as bench.py does."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
    from pn2_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.scene_ids_for_rank(10, rank, world)          # strong split of a fixed job
    weak = sharding.weak_scene_ids(rank, batch=3, step=2)       # bench.py's weak-scaling ids
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    mx = sharding.max_over_ranks(t)
    total = sharding.sum_over_ranks(torch.tensor([float(len(ids))], dtype=torch.float64))
    q.put((rank, ids, weak, float(mx), float(total)))
    dist.destroy_process_group()


def test_scene_sharding_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, weak0, mx0, tot0), (r1, ids1, weak1, mx1, tot1) = res
    assert sorted(ids0 + ids1) == list(range(10)) and not set(ids0) & set(ids1)   # every scene exactly once
    assert abs(len(ids0) - len(ids1)) <= 1                                         # balanced
    assert not set(weak0) & set(weak1) and len(weak0) == len(weak1) == 3            # disjoint, equal per-rank work
    assert mx0 == mx1 == 2.0 and tot0 == tot1 == 10.0


def test_sharding_single_process_is_identity():
    sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
    from pn2_b200 import sharding
    assert sharding.scene_ids_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    t = torch.tensor([3.0])
    assert float(sharding.max_over_ranks(t)) == 3.0
