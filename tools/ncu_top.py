"""Top stall lines of an ncu --page source --csv dump: python tools/ncu_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
i_src, i_smp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > i_smp and r[i_smp].isdigit()]
total = sum(int(r[i_smp]) for r in body)
print("total samples", total, "instructions", len(body))
ops = {}
for r in body:
    op = r[i_src].split()[0] if r[i_src].split() else "?"
    if op.startswith("@"):
        op = r[i_src].split()[1]
    ops.setdefault(op, [0, 0])
    ops[op][0] += int(r[i_smp])
    ops[op][1] += int(r[i_exec])
print("--- by opcode (samples, executed)")
for op, (s, e) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{op:28s} {s:7d} {100.0 * s / total:5.1f}%  exec={e}")
print("--- top lines")
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][i_smp]))[:n]:
    print(f"{k:5d} {int(r[i_smp]):6d} {100.0 * int(r[i_smp]) / total:5.1f}%  {r[i_src].strip()[:110]}")
