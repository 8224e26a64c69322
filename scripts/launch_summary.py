#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel the launches, total and mean device time,
and the share of the listed launches (cold-cache, serialised times: compare SHARES, not absolutes).
    python scripts/launch_summary.py gpurun_out/r02_launches_bench_c4.csv > profiles/r02_launches_bench_c4.md"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
order = []
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    order.append((name, ms))
total = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | mean ms | share |")
print("|---|---|---|---|---|")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.4f | %.1f %% |" % (name, n, ms, ms / n, 100 * ms / total))
print("\n%d launches, %.2f ms in total" % (sum(a[0] for a in agg.values()), total))
