"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals over a window of the run.
python scripts/launch_summary.py file.csv [lo_frac hi_frac]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.45, 0.75)
seg = rows[int(len(rows) * lo):int(len(rows) * hi)]
d, c = defaultdict(float), defaultdict(int)
for r in seg:
    name = r[4].split("(")[0][-64:]
    d[name] += float(r[-1].replace(",", ""))
    c[name] += 1
tot = sum(d.values())
print(f"{len(rows)} launches in the file; window {lo}-{hi}: {len(seg)} launches, {tot / 1000:.1f} us")
for k, v in sorted(d.items(), key=lambda x: -x[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print(f"{v / 1000:9.1f} us {100 * v / tot:5.1f}%  x{c[k]:<3d} {v / 1000 / c[k]:8.1f} us each  {k}")
