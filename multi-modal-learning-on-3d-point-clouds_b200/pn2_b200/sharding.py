"""Scene sharding across GPUs (one process per GPU).  Every operator on the path is per-cloud, so scenes are
partitioned round-robin over the ranks, weights are replicated and there is NO collective on the forward path
(the reference's only parallelism is torch.nn.DataParallel, train_scannet_semseg.py:100-106).
The helpers below are the only cross-rank traffic bench.py issues: reductions of timing scalars."""
import torch
import torch.distributed as dist


def scene_ids_for_rank(num_scenes, rank, world):
    """Round-robin split of a fixed job of `num_scenes` scenes."""
    return list(range(rank, num_scenes, world))


def weak_scene_ids(rank, batch, step):
    """Ids of the `batch` scenes rank `rank` processes at step `step` (weak scaling: fixed work per rank)."""
    base = 100000 * rank + step * batch
    return list(range(base, base + batch))


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def max_over_ranks(t):
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def sum_over_ranks(t):
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
