"""Per-kernel durations of the LAST graph replay / step in an `ncu --metrics gpu__time_duration.sum --csv` log:
    python profiles/launch_list.py <log.csv> [first kernel name fragment = match_kernel]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
data = [(r[ki], float(r[mi].replace(",", "")) * scale.get(r[ui], 1e-3)) for r in rows[h + 1:] if len(r) > mi]
frag = sys.argv[2] if len(sys.argv) > 2 else "match_kernel"
# a step starts with the two matcher launches and ends with finalize_kernel; the log also holds warm-up steps, the
# profiled passes, matcher-only passes and the end-to-end passes: take the SECOND complete step (first timed one after
# one warm-up step)
starts = [i for i, (n, _) in enumerate(data) if frag in n]
steps = []
for a in starts:
    seg = []
    for n, v in data[a:]:
        seg.append((n, v))
        if "finalize_kernel" in n:
            break
    else:
        continue
    if sum(1 for n, _ in seg if frag in n) == 2 and any("ap_kernel" in n for n, _ in seg):
        steps.append(seg)
seg = [d for d in steps[min(1, len(steps) - 1)] if "FillFunctor" not in d[0]]
agg = collections.OrderedDict()
for n, v in seg:
    n = n.split("(")[0].replace("void ", "").replace("orie::", "").replace("<unnamed>::", "")
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(seg)} launches, {tot:.1f} us summed kernel time (ncu: serialised, cold caches)")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v / tot * 100:5.1f}%  {v:9.1f} us  x{c:3d}  {n}")
