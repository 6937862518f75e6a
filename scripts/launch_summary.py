"""Summarises an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): launches and total time per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("upc::", "")[:64]
    t = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t * scale
tot = sum(v[1] for v in agg.values())
print("%-66s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %6d %12.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("%-66s %6d %12.1f" % ("all", sum(v[0] for v in agg.values()), tot))
