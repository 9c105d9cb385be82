#!/usr/bin/env python
"""Per-opcode / per-instruction stall summary from an ncu source page.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:k_forward --launch-count 1 > /tmp/src.csv
    python profiles/stalls.py /tmp/src.csv [top_n]
"""
import collections
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
print("#", rows[0][1])
h = rows[1]
si, ai, ii = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) >= len(h)]
tot = sum(num(r[ai]) for r in data)
print(f"samples {tot}, SASS instructions {len(data)}, warp instructions executed {sum(num(r[ii]) for r in data)}")
agg = collections.Counter()
for r in data:
    for i, c in stalls:
        agg[c] += num(r[i])
print("stall reasons:", ", ".join(f"{c[6:]} {v / max(1, tot) * 100:.1f}%" for c, v in agg.most_common(9)))
cls, inst = collections.Counter(), collections.Counter()
for r in data:
    t = r[si].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    cls[op] += num(r[ai])
    inst[op] += num(r[ii])
print("\n| opcode | samples | % | warp instr |\n|---|---:|---:|---:|")
for op, c in cls.most_common(16):
    print(f"| {op} | {c} | {c / max(1, tot) * 100:.1f} | {inst[op]} |")
print("\ntop instructions by samples:")
for r in sorted(data, key=lambda r: -num(r[ai]))[:topn]:
    why = max(stalls, key=lambda ic: num(r[ic[0]]))[1][6:]
    print(f"{num(r[ai]):6d}  {why:10s} {r[si].strip()[:100]}")
