"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel markdown table.
usage: python tools/launch_summary.py <launches.csv> > summary.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ci["Kernel Name"]]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    a = agg.setdefault(name, dict(n=0, us=0.0, grid=r[ci["Grid Size"]], block=r[ci["Block Size"]]))
    a["n"] += 1
    a["us"] += us
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches | total ms | share | avg us | grid | block |")
print("|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print("| `%s` | %d | %.3f | %.1f%% | %.1f | %s | %s |" % (name[:75], a["n"], a["us"] / 1e3, 100 * a["us"] / tot, a["us"] / a["n"], a["grid"], a["block"]))
print("\ntotal %.3f ms over %d launches" % (tot / 1e3, sum(a["n"] for a in agg.values())))
