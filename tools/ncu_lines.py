"""Per-source-line stall samples and executed instructions from `ncu -i rep --page source --csv --print-source cuda,sass > file`.
    python tools/ncu_lines.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg = {}
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) < 10 or not r[0].isdigit():
        continue
    k = (fname, int(r[0]), r[1])
    a = agg.setdefault(k, [0, 0])
    a[0] += num(r[6])
    a[1] += num(r[7])
print("samples", sum(v[0] for v in agg.values()), "warp instructions", sum(v[1] for v in agg.values()))
for title, idx in (("by stall samples", 0), ("by executed instructions", 1)):
    print("---", title)
    for k in sorted(agg, key=lambda k: -agg[k][idx])[:n]:
        print("%-18s %4d %6d %9d  %s" % (k[0][:18], k[1], agg[k][0], agg[k][1], k[2].strip()[:110]))
