"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn]
        short = name.split("(")[0][-70:] if not name.startswith("void ob::") and "ob::" not in name else name.split("(")[0]
        agg.setdefault(short, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{sum(len(v) for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):5d} x avg {sum(v) / len(v) / 1e3:9.1f} us  total {sum(v) / 1e6:8.3f} ms  {100 * sum(v) / tot:5.1f}%  {k}")
