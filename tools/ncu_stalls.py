#!/usr/bin/env python
"""Aggregate the ncu source page (SASS view) of one kernel: total stall samples by reason and the hottest
instruction windows.   python tools/ncu_stalls.py <source.csv> [window]"""
import csv
import sys

allrows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[which]:starts[which + 1]]
print(rows[0][1])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
si = hdr.index("# Samples")
ii = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print(f"{len(data)} SASS instructions, {tot} samples, {sum(int(r[ii] or 0) for r in data)} warp instructions executed")
agg = {h: sum(int(r[i] or 0) for r in data) for i, h in stall_cols}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f"  {h:28s} {v:8d} {v / max(tot, 1):6.1%}")
print(f"hottest windows of {win} instructions:")
wins = []
for s in range(0, len(data), win):
    chunk = data[s:s + win]
    wins.append((sum(int(r[si] or 0) for r in chunk), s))
for v, s in sorted(wins, reverse=True)[:12]:
    chunk = data[s:s + win]
    ops = {}
    for r in chunk:
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[si] or 0)
    top = ", ".join(f"{k}:{n}" for k, n in sorted(ops.items(), key=lambda x: -x[1])[:6])
    reasons = {h: sum(int(r[i] or 0) for r in chunk) for i, h in stall_cols}
    rs = ", ".join(f"{k[6:]}:{n}" for k, n in sorted(reasons.items(), key=lambda x: -x[1])[:4])
    print(f"  inst {s:6d}-{s + win:6d}: {v:7d} samples ({v / max(tot, 1):5.1%}) exec/inst {int(chunk[0][ii] or 0):9d} | {top} | {rs}")
