"""Kernel share of a bench command from the ncu launch list (--metrics gpu__time_duration.sum --csv):
    python tools/launch_summary.py profiles/r01_bench_launches.csv > profiles/r01_bench_launches_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[rows.index(hdr) + 1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    t = tot[r[ik]]
    t[0] += 1
    t[1] += ms
total = sum(t[1] for t in tot.values())
print("kernel share of the bench command (ncu --metrics gpu__time_duration.sum, %d launches, %.1f ms total)" % (
    sum(t[0] for t in tot.values()), total))
for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    name = k.split("(")[0][-42:]
    print("%-42s launches %4d  total %9.3f ms  share %5.1f%%" % (name, t[0], t[1], 100 * t[1] / total))
