#!/usr/bin/env python
"""Per-source-line instruction counts of one loop of a kernel (needs -lineinfo).
usage: sass_lines.py <obj.o|lib.so> <function-substring> [loop-index]   (loops sorted by size, >200 instructions)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, key = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
txt = ""
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        txt += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
on, cur, ins, labels, pending = False, ("?", 0), [], {}, []
for ln in txt.splitlines():
    if ".section" in ln and ".text." in ln:
        on = key in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"(\.L_x_\d+):", ln)
    if m:
        pending.append(m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*(.*?);", ln)
    if m:
        for lb in pending:
            labels[lb] = int(m.group(1), 16)
        pending = []
        ins.append((int(m.group(1), 16), m.group(3), m.group(5), cur))
loops = []
for a, op, rest, _ in ins:
    if op == "BRA":
        m = re.search(r"(\.L_x_\d+)", rest)
        if m and labels.get(m.group(1), 1 << 60) < a:
            loops.append((labels[m.group(1)], a))
loops = [l for l in sorted(loops, key=lambda x: x[1] - x[0]) if sum(1 for i in ins if l[0] <= i[0] <= l[1]) > 200]
lo, hi = loops[which]
body = [i for i in ins if lo <= i[0] <= hi]
print("loop 0x%x..0x%x: %d instructions" % (lo, hi, len(body)))
c = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for a, op, rest, src in body:
    c[src] += 1
    ops[src][op] += 1
for src, n in sorted(c.items()):
    print("%-24s %5d   %s" % ("%s:%d" % src, n, " ".join("%s:%d" % kv for kv in ops[src].most_common(6))))
