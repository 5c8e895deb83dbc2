"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one bench step."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[mi].replace(",", ""))) for r in rows[h + 1:] if len(r) > mi]
starts = [i for i, (n, _) in enumerate(data) if "match_kernel" in n]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # which step (0 = first warm-up)
a, b = starts[2 * which], starts[2 * which + 2]
agg = collections.OrderedDict()
for n, v in data[a:b]:
    n = n.split("(")[0].replace("void ", "").replace("orie::", "")
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"step {which}: {b - a} launches, {tot / 1e6:.3f} ms summed kernel time (ncu: cold cache, serialised)")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v / tot * 100:5.1f}%  {v / 1e3:9.1f} us  x{c:3d}  {n}")
