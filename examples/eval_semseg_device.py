"""Evaluation loop of train_scannet_semseg.py:180-250 with everything on the device (synthetic scenes):
pinned host batch -> H2D -> predict() (arg-max fused into the head kernel) -> point-wise and voxel-wise counters
(pn2_b200.pc_util) -> one small read-back at the end.  The reference brings logits, targets, weights and coordinates to
the host every batch and voxelises scene by scene in numpy.

    python examples/eval_semseg_device.py [--batches 20] [--batch 32] [--check]

--check also runs the reference's host-side loop (restated in oracle/pc_util_ref.py) on the same predictions.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-learning-on-3d-point-clouds_b200"))
from pn2_b200 import pc_util, scenes  # noqa: E402
from pn2_b200.models import PipelinedForward, PointNet2SemSeg  # noqa: E402

NUM_CLASSES, NPOINTS = 21, 8192


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    from pn2_b200 import pointnet_util
    pointnet_util.set_mlp_precision("bf16")  # tensor-core shared MLP (2e-2); the modules' default is fp32
    model = PointNet2SemSeg(NUM_CLASSES).eval().to(dev)
    B = args.batch
    rng = np.random.default_rng(0)
    hosts = []
    for i in range(min(args.batches, 8)):
        pts = torch.from_numpy(scenes.scannet_batch(4000 + i * B, B, NPOINTS)).pin_memory()             # (B, N, 6)
        target = torch.from_numpy(rng.integers(0, NUM_CLASSES, (B, NPOINTS)).astype(np.int64)).pin_memory()
        weights = torch.from_numpy((rng.random((B, NPOINTS)) < 0.9).astype(np.float32)).pin_memory()
        hosts.append((pts, target, weights))
    ex = hosts[0][0].to(dev).permute(0, 2, 1)
    pipe = PipelinedForward(model, ex[:, :3].contiguous(), ex[:, 3:].contiguous(), depth=args.depth, labels=True)
    counters = pc_util.EvalCounters(NUM_CLASSES, dev, res=0.02)
    preds = []
    # one staging set per pipeline slot; copies run on their own stream and overlap the forwards of the other slots
    depth = args.depth
    stage = [(torch.empty((B, NPOINTS, 6), device=dev), torch.empty((B, NPOINTS), dtype=torch.int64, device=dev),
              torch.empty((B, NPOINTS), device=dev)) for _ in range(depth)]
    slot_done = [None] * depth
    copy_in = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(args.batches):
            k = i % depth
            pts, target, weights = stage[k]
            with torch.cuda.stream(copy_in):
                if slot_done[k] is not None:
                    copy_in.wait_event(slot_done[k])  # the slot's previous forward + metric have consumed its buffers
                for dst, src in zip(stage[k], hosts[i % len(hosts)]):
                    dst.copy_(src, non_blocking=True)
                h2d = torch.cuda.Event()
                h2d.record(copy_in)
            x = pts.permute(0, 2, 1)
            labels, _, st = pipe.submit(x[:, :3], x[:, 3:], after=h2d)
            with torch.cuda.stream(st):  # the metric follows the forward on the slot's stream
                counters.update(pts, target, labels, weights)
                if args.check:
                    preds.append(labels.long().cpu().numpy())
                slot_done[k] = torch.cuda.Event()
                slot_done[k].record(st)
        pipe.join()
    res = counters.result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d scenes in %.3f s -> %.0f scenes/s including the metric; point acc %.4f, voxel acc %.4f" %
          (args.batches * B, dt, args.batches * B / dt, res["total_correct"] / max(res["total_seen"], 1),
           res["total_correct_vox"] / max(res["total_seen_vox"], 1)))
    if args.check:
        from oracle import pc_util_ref
        t0 = time.perf_counter()
        want = None
        for i in range(args.batches):
            pts, target, weights = (t.numpy() for t in hosts[i % len(hosts)])
            c = pc_util_ref.voxel_accuracy_counts(pts, target, preds[i], weights, NUM_CLASSES, res=0.02)
            want = c if want is None else {k: want[k] + np.asarray(v) for k, v in c.items()}
        dt_ref = time.perf_counter() - t0
        for k, v in want.items():
            assert np.array_equal(np.asarray(res[k]), np.asarray(v)), k
        print("host-side voxel metric of the reference loop on the same predictions: identical counters, %.3f s (%.0f scenes/s)" %
              (dt_ref, args.batches * B / dt_ref))


if __name__ == "__main__":
    main()
