"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`).

    python scripts/summarize_launches.py launches.csv "header line" > summary.txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
n = 0
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("pn2::<unnamed>::", "").replace("(anonymous namespace)::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    n += 1
total = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("%d launches, %.1f us total (cold-cache, serialised: compare shares)\n" % (n, total))
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-72s n=%3d %10.1f us  each %8.1f us %5.1f%%" % (name[:72], c, us, us / c, 100 * us / total))
