"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
warp instructions executed and stall samples per line (top N), plus totals per line range.
usage: ncu_lines.py export.csv [topN] [lo-hi ...]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ranges = [tuple(int(x) for x in a.split("-")) for a in sys.argv[3:]]
rows = list(csv.reader(open(path)))
hdr = None
per = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue  # keep the per-line summary rows (Address == "-")
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    inst = int(d.get("Instructions Executed", "0") or 0)
    samp = int(d.get("# Samples", "0") or 0)
    e = per.setdefault(line, [0, 0, r[1]])
    e[0] += inst
    e[1] += samp
tot_i = sum(v[0] for v in per.values())
tot_s = sum(v[1] for v in per.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for lo, hi in ranges:
    i = sum(v[0] for k, v in per.items() if lo <= k <= hi)
    s = sum(v[1] for k, v in per.items() if lo <= k <= hi)
    print("lines %4d-%4d: inst %5.1f %%  samples %5.1f %%" % (lo, hi, 100.0 * i / max(tot_i, 1), 100.0 * s / max(tot_s, 1)))
print("-- top lines by samples")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5d inst %5.1f %% samp %5.1f %%  %s" % (k, 100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1), v[2].strip()[:110]))
