"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (second half = the measured pass)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
data = [r for r in rows[hi + 1:] if len(r) > mv]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
half = data[int(len(data) * frac):]
agg = collections.OrderedDict()
for r in half:
    name = r[kn].split("(")[0].replace("void ", "").replace("seald::", "")
    v = float(r[mv].replace(",", ""))
    us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[mu], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% |" % (k[:70], c, t, 100 * t / tot))
print("| **total** | %d | %.1f | |" % (len(half), tot))
