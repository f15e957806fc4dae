#!/usr/bin/env python
"""Instruction mix of the fast-path loop of a k_fields instantiation (reads cuobjdump -sass).
usage: sass_mix.py <lib.so> <mangled-substring>   e.g. IfLb1ELi9ELb0EEE"""
import collections
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
lines, on = [], False
for ln in txt.splitlines():
    if "Function :" in ln:
        on = ("k_fields" in ln and key in ln)
    elif on:
        lines.append(ln)
ins = []
for ln in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(5)))
# the fast loop is the region containing the first 16 SHFL.UP
shfl = [a for a, op, mod, rest in ins if op == "SHFL"]
print("total instructions", len(ins), "bytes", ins[-1][0] + 16)
# find backward branch enclosing first SHFL
first = shfl[0]
best = None
for a, op, mod, rest in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m:
            t = int(m.group(1), 16)
            if t < a and t <= first <= a:
                if best is None or (a - t) < (best[1] - best[0]):
                    best = (t, a)
print("loop containing first SHFL: 0x%x..0x%x" % best)
# split the loop at the midpoint between the two SHFL groups if both variants are inside
lo, hi = best
body = [(a, op, mod) for a, op, mod, rest in ins if lo <= a <= hi]
grp1 = [a for a in shfl if a < shfl[0] + 0x400]
c = collections.Counter()
# fast variant = the contiguous region around the first SHFL group up to the jump that skips the other variant
print("loop body instrs", len(body))
for a, op, mod in body:
    key2 = op
    if op in ("FRND", "F2I", "I2F", "I2FP", "IMAD", "SHF", "FADD", "FFMA", "FMUL"):
        key2 = op + (mod if op in ("FRND", "F2I", "I2F") else (".RZ" if ".RZ" in mod else ".RM" if ".RM" in mod else (".HI" if ".HI" in mod else ".MOV" if ".MOV" in mod else "")))
    c[key2] += 1
for k, v in c.most_common(40):
    print("%-14s %6d" % (k, v))
