"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/ncu_launches.py file.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= i_val:
        continue
    v = float(r[i_val].replace(",", ""))
    v = v / 1000.0 if r[i_unit] in ("ns", "nsecond") else v  # -> us
    name = r[i_name].split("(")[0].replace("edm::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"{'kernel':48s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>9s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} {n:8d} {v:12.1f} {100 * v / tot:6.1f}% {v / n:9.1f}")
print(f"{'total':48s} {sum(n for n, _ in agg.values()):8d} {tot:12.1f}")
