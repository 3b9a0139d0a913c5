"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    t = float(r[vi].replace(",", ""))
    t = t / 1e3 if r[ui] in ("ns", "nsecond") else t * 1e3 if r[ui] in ("ms", "msecond") else t
    name = re.sub(r"\(.*", "", r[ki])[:110]
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% n={v[0]:4d} avg={v[1] / v[0]:8.1f}  {k}")
