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


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
    from pn2_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # replicated weights
    net = torch.nn.Sequential(torch.nn.Conv1d(6, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
    grads = sharding.FlatGradients(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)  # every rank its own shard of scenes
    local, after = [], None
    for step in range(2):
        x = torch.randn(4, 6, 32, generator=g)
        grads.zero()
        net(x).square().mean().backward()  # accumulates in place into the views of the flat buffer
        local.append(grads.flat.clone())
        grads.all_reduce_mean()
        if step == 0:
            after = grads.flat.clone()
        opt.step()
    q.put((rank, [t.numpy() for t in local], after.numpy(), torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy(),
           [p.grad.data_ptr() == grads.flat[o:o + 1].data_ptr() for p, o in zip(grads.params, np.cumsum([0] + [p.numel() for p in grads.params])[:-1])]))
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_two_ranks_gloo():
    """Training exchange: the gradients of all parameters live in one flat buffer that is averaged by ONE
    all-reduce per step; afterwards both ranks hold the mean of the per-rank gradients and identical weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, loc0, after0, w0, views0), (_, loc1, after1, w1, views1) = res
    assert all(views0) and all(views1)                                  # .grad tensors are views into the flat buffer
    assert not np.array_equal(loc0[0], loc1[0])                         # different shards, different local gradients
    np.testing.assert_allclose(after0, (loc0[0] + loc1[0]) / 2, rtol=0, atol=1e-7)
    np.testing.assert_array_equal(after0, after1)                       # both ranks hold the same averaged gradient
    np.testing.assert_array_equal(w0, w1)                               # and stay in lock-step after two updates
