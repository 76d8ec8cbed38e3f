"""Aggregate the warp-stall samples of an exported ncu source page (ncu -i rep --page source --csv --print-source
cuda,sass) per (file, source line), across all files of the kernel. Usage: python tools/ncu_lines2.py export.csv [top]"""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
names = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_membar", "stall_math",
         "stall_mio", "stall_lg", "stall_branch_resolving", "stall_no_inst", "stall_selected"]
per = collections.defaultdict(lambda: [0.0] * (2 + len(names)))
text = {}
fname, hdr, cur = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        isamp, iexec = hdr.index("# Samples"), hdr.index("Instructions Executed")
        idx = [hdr.index(n) for n in names]
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0].strip().isdigit():
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()[:90]
        continue
    if len(r) > 3 and r[2].startswith("0x"):
        def f(i):
            try:
                return float(r[i])
            except (ValueError, IndexError):
                return 0.0
        d = per[cur]
        d[0] += f(isamp)
        d[1] += f(iexec)
        for k, i in enumerate(idx):
            d[2 + k] += f(i)
tot = sum(d[0] for d in per.values())
print("total samples", tot, " warp instructions", sum(d[1] for d in per.values()))
print("%-28s samples   pct   executed | %s" % ("file:line", " ".join(n.replace("stall_", "")[:7] for n in names)))
for key, d in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %5d %7.0f %5.1f%% %10.0f | %s | %s" % (key[0][:22], key[1], d[0], 100 * d[0] / tot, d[1],
                                                       " ".join("%7.0f" % v for v in d[2:]), text.get(key, "")))
