"""Aggregate the warp-stall samples of an exported ncu source page (--page source --csv --print-source cuda,sass)
per source line. Usage: python tools/ncu_lines.py export.csv source.cu [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
isamp = hdr.index("# Samples")
iexec = hdr.index("Instructions Executed")
names = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_membar", "stall_math", "stall_selected"]
idx = [hdr.index(n) for n in names]
src = open(sys.argv[2]).read().split("\n")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
per = collections.defaultdict(lambda: [0.0] * (2 + len(names)))
cur = None
for r in rows[hi + 1:]:
    if r and r[0].strip().isdigit():
        cur = int(r[0])
        continue
    if len(r) == len(hdr) and r[2].startswith("0x"):
        def f(i):
            try:
                return float(r[i])
            except ValueError:
                return 0.0
        d = per[cur]
        d[0] += f(isamp)
        d[1] += f(iexec)
        for k, i in enumerate(idx):
            d[2 + k] += f(i)
tot = sum(d[0] for d in per.values())
print("total samples", tot, " warp instructions", sum(d[1] for d in per.values()))
print(" line  samples   pct   executed |", " ".join(n.replace("stall_", "") for n in names))
for ln, d in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[ln - 1].strip()[:80] if ln and ln <= len(src) else ""
    print("%5d %7.0f %5.1f%% %10.0f | %s | %s" % (ln, d[0], 100 * d[0] / tot, d[1], " ".join("%6.0f" % v for v in d[2:]), text))
