#!/usr/bin/env python
"""Join the SASS source page of an ncu report with nvdisasm line info: stall samples (barrier waits excluded) and
executed warp instructions per CUDA source line.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; cuobjdump -xelf all obj.o ; nvdisasm -g x.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt <mangled kernel name> [top]"""
import collections
import csv
import re
import sys

src_csv, dis_txt, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.lstrip().startswith(".section") and ".text." + kernel in l)
cur, off2line = None, {}
for l in lines[start + 1:]:
    if l.lstrip().startswith(".section"):
        break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        if "inlined at" not in m.group(3) or cur is None:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*)", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ai, si, ii, bi = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_barrier")
addr = lambda r: int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])  # noqa: E731
base = addr(data[0])
agg, exe = collections.Counter(), collections.Counter()
for r in data:
    ln = off2line.get(addr(r) - base)
    agg[ln] += int(r[si] or 0) - int(r[bi] or 0)
    exe[ln] += int(r[ii] or 0)
tot, tote = sum(agg.values()), sum(exe.values())
print(f"{tot} non-barrier samples, {tote} warp instructions")
for ln, v in agg.most_common(top):
    print(f"{str(ln):36s} samples {v:8d} {v / tot:6.1%}   instructions {exe[ln]:12d} {exe[ln] / tote:6.1%}")
