#!/usr/bin/env python
"""Split an ncu SASS source page (--page source --csv) of one kernel at its BAR.SYNC instructions and print,
per phase, the warp instructions executed, stall samples and the opcode mix.
    python profiles/phases.py page.csv"""
import collections
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
si, ai, ii = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
seen, data = set(), []
for r in rows[2:]:
    if len(r) >= len(h) and r[0] not in seen:      # some exports list every row twice
        seen.add(r[0])
        data.append(r)
phases, cur = [], []
for r in data:
    cur.append(r)
    if "BAR.SYNC" in r[si]:
        phases.append(cur)
        cur = []
phases.append(cur)
tot_i = sum(num(r[ii]) for r in data)
tot_s = sum(num(r[ai]) for r in data)
print(f"# {rows[0][1]}: {tot_i} warp instructions, {tot_s} samples, {len(phases)} phases (split at BAR.SYNC)")
for k, ph in enumerate(phases):
    ins, smp = sum(num(r[ii]) for r in ph), sum(num(r[ai]) for r in ph)
    ops = collections.Counter()
    for r in ph:
        t = r[si].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += num(r[ii])
    mix = ", ".join(f"{o} {c / max(1, ins) * 100:.0f}%" for o, c in ops.most_common(7))
    print(f"phase {k}: {len(ph)} SASS, instr {ins / max(1, tot_i) * 100:.1f}%, samples {smp / max(1, tot_s) * 100:.1f}% | {mix}")
