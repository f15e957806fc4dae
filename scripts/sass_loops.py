#!/usr/bin/env python
"""Instruction mix of every innermost loop of a kernel (reads cuobjdump -sass).
usage: sass_loops.py <lib.so|obj.o> <function-substring>"""
import collections
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
lines, on = [], False
for ln in txt.splitlines():
    if "Function :" in ln:
        on = key in ln
    elif on:
        lines.append(ln)
ins = []
for ln in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(5)))
print("total instructions", len(ins))
loops = []
for a, op, mod, rest in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m:
            t = int(m.group(1), 16)
            if t < a:
                loops.append((t, a))
for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
    body = [(a, op, mod) for a, op, mod, rest in ins if lo <= a <= hi]
    if len(body) < 200:
        continue
    c = collections.Counter()
    for a, op, mod in body:
        k = op
        if op in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL"):
            k = op + (".U8" if ".U8" in mod else (".64" if ".64" in mod else (".128" if ".128" in mod else "")))
        c[k] += 1
    print("loop 0x%x..0x%x: %d instructions" % (lo, hi, len(body)))
    print("  ", ", ".join("%s %d" % kv for kv in c.most_common(28)))
