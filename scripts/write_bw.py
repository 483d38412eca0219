"""Write-only, read-only and copy bandwidth of the device (what bounds a write-dominated operator such as three_interpolate)."""
import json, torch
dev = torch.device("cuda:0")
n = 1 << 28  # 1 GiB of fp32
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)


def t_ms(fn, iters=10):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2]


out = {}
out["fill_gbs"] = 4 * n / t_ms(lambda: a.fill_(1.0)) / 1e6
out["zero_gbs"] = 4 * n / t_ms(lambda: a.zero_()) / 1e6
out["copy_gbs_read_plus_write"] = 8 * n / t_ms(lambda: b.copy_(a)) / 1e6
out["sum_read_gbs"] = 4 * n / t_ms(lambda: a.sum()) / 1e6
# 128 MiB writes (the size of three_interpolate's output at the C1 shape), L2-sized effects included
c = torch.empty(1 << 25, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(9):
    flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); c.fill_(1.0); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts.sort()
out["fill_128MiB_after_flush_gbs"] = 4 * (1 << 25) / ts[len(ts) // 2] / 1e6
out["fill_128MiB_after_flush_ms"] = ts[len(ts) // 2]
print(json.dumps(out))
