"""Host-link probe: pinned D2H / H2D bandwidth per rank and in aggregate when N ranks copy at once (one process per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/d2h_probe.py

The end-to-end leg of bench.py reads the (32, 8192, 21) fp32 logits back every step: 22 MB per step per GPU.  This probe
measures what the box's host side sustains for exactly that transfer pattern (22 MB chunks, pinned memory, one copy stream
per direction), so the e2e scaling can be judged against the box's ceiling.  Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

CHUNK = 32 * 8192 * 21 * 4


def numa_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        devid = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devid)
        return int(open(path).read().strip())
    except Exception:  # noqa: BLE001
        return None


def cpus_of_node(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        out = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:  # noqa: BLE001
        return None


def measure(device, direction, seconds=0.6, streams=1):
    dev_bufs = [torch.empty(CHUNK, dtype=torch.uint8, device=device) for _ in range(2 * streams)]
    host_bufs = [torch.empty(CHUNK, dtype=torch.uint8).pin_memory() for _ in range(2 * streams)]
    sts = [torch.cuda.Stream(device) for _ in range(2 * streams)]
    torch.cuda.synchronize()
    n = 0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in sts:
        s_.wait_event(e0)
    while time.perf_counter() - t0 < seconds:
        for k in range(streams):
            if direction in ("d2h", "both"):
                with torch.cuda.stream(sts[2 * k]):
                    host_bufs[2 * k].copy_(dev_bufs[2 * k], non_blocking=True)
            if direction in ("h2d", "both"):
                with torch.cuda.stream(sts[2 * k + 1]):
                    dev_bufs[2 * k + 1].copy_(host_bufs[2 * k + 1], non_blocking=True)
        n += streams
        if n % 8 == 0:
            for s_ in sts:
                s_.synchronize()
    for s_ in sts:
        torch.cuda.current_stream().wait_stream(s_)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    per_dir = n * CHUNK / ms / 1e6  # GB/s in each active direction
    return per_dir


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    node = numa_of_gpu(local)
    res = {"rank": rank, "gpu_numa_node": node, "affinity": len(os.sched_getaffinity(0))}
    for bind in (False, True):
        if bind:
            cpus = cpus_of_node(node) if node is not None and node >= 0 else None
            if not cpus:
                res["bound"] = "no NUMA information for this GPU; not re-bound"
                break
            os.sched_setaffinity(0, cpus)
        for direction in ("d2h", "h2d", "both"):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            res["%s%s" % (direction, "_bound" if bind else "")] = measure(device, direction)
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        keys = [k for k in allres[0] if k not in ("rank", "gpu_numa_node", "affinity", "bound")]
        agg = {k: sum(r.get(k, 0.0) for r in allres) for k in keys}
        print(json.dumps({"n_gpus": world, "chunk_bytes": CHUNK, "aggregate_gbs": agg, "per_rank": allres,
                          "host_cpus": os.cpu_count(),
                          "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
                          if os.path.isdir("/sys/devices/system/node") else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
