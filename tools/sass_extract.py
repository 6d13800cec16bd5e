#!/usr/bin/env python
"""Trimmed SASS evidence of the hot kernels: opcode histogram and the hottest straight-line window of each.

    python tools/sass_extract.py > profiles/r02_sass_hot_loops.txt

Reads the object files of the in-tree build (csrc/build/*.o) with cuobjdump; no GPU needed."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tensor-networks-simple-update_b200", "csrc", "build")
# (object file, mangled-name regex, label, opcode whose densest window is shown, window length)
KERNELS = [
    ("lqc_inst.o", r"lq_chol_warp_kernelILi2ELb0E", "K1c lq_chol_warp_kernel<2, real>: streaming CholeskyQR2", "DMMA", 56),
    ("lqc_inst.o", r"lq_chol_warp_kernelILi4ELb1E", "K1c lq_chol_warp_kernel<4, complex128> (real embedding)", "DMMA", 40),
    ("svdw_inst.o", r"svd_warp_kernelILi16ELb0E", "K2 svd_warp_kernel<16, V-free>: register-resident Jacobi", "DFMA", 72),
    ("lqc_inst.o", r"recompose_chol_kernelILi2ELb0E", "K3c recompose_chol_kernel<2, real>: in-place recompose", "DMMA", 56),
    ("engine.o", r"pair_rdm_dmma_kernelILi2E", "pair_rdm_dmma_kernel<2>: Gram matrices of energy_per_site", "DMMA", 48),
    ("fused_inst.o", r"run_fused_kernelILi8ELi8ELi2ELb0E", "run_fused_kernel<8, 8, 2, V-free>: on-chip run() (config 3 shape)", "DFMA", 48),
]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, funcs = None, collections.OrderedDict()
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def opcode(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    return op


cache = {}
for obj, pat, label, hot, win in KERNELS:
    if obj not in cache:
        cache[obj] = functions(obj)
    name = next((n for n in cache[obj] if re.search(pat, n)), None)
    if name is None:
        print(f"== {label}: not found in {obj}\n")
        continue
    ins = cache[obj][name]
    hist = collections.Counter(opcode(t).split(".")[0] for _, t in ins)
    full = collections.Counter(opcode(t) for _, t in ins)
    print(f"== {label}\n   {name}\n   {len(ins)} SASS instructions; "
          + ", ".join(f"{k} {v}" for k, v in hist.most_common(12)))
    wide = {k: v for k, v in full.items() if re.search(r"\.(128|256)|DMMA|UBLKCP|REDUX|MUFU\.RSQ64H|UTC|LDTM", k)}
    print("   wide / tensor / special: " + (", ".join(f"{k} {v}" for k, v in sorted(wide.items(), key=lambda x: -x[1])) or "-"))
    hits = [i for i, (_, t) in enumerate(ins) if opcode(t).startswith(hot)]
    if hits:
        best, best_n = 0, -1
        for i in range(0, max(1, len(ins) - win)):
            n = sum(1 for h in hits if i <= h < i + win)
            if n > best_n:
                best, best_n = i, n
        print(f"   densest {win}-instruction window for {hot} ({best_n} of them):")
        for addr, t in ins[best:best + win]:
            print(f"      /*{addr:05x}*/  {t}")
    print()
