"""Scene sharding across GPUs (one process per GPU).  Every operator on the path is per-cloud, so scenes are
partitioned round-robin over the ranks, weights are replicated and there is NO collective on the forward path
(the reference's only parallelism is torch.nn.DataParallel, train_scannet_semseg.py:100-106).
The helpers below are the only cross-rank traffic of the path: reductions of timing scalars (bench.py) and, for training,
the all-reduce of one flat gradient buffer (FlatGradients)."""
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


class FlatGradients:
    """All parameter gradients of a module as views into ONE flat fp32 buffer, so that the gradient exchange of scene-sharded
    training is a single all-reduce (the only collective of the path; the reference uses DataParallel's per-tensor
    reduction, train_scannet_semseg.py:100-106).  `zero()` clears the buffer (backward accumulates in place into the views),
    `all_reduce_mean()` averages it over the ranks; without an initialised process group both degrade to the single-process
    behaviour.  Used by pn2_b200.models.GraphedTrainStep between its backward graph and its optimizer graph."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        if not self.params:
            raise ValueError("FlatGradients needs at least one parameter that requires a gradient")
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradients handles fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        w = self.world
        if w > 1:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(w)
        return self.flat
