#!/usr/bin/env python
"""SASS of a kernel with the scheduling control fields decoded (stall count, yield, write/read barrier, wait mask).
usage: sass_ctrl.py <obj> <function-substring> [start-hex end-hex]"""
import re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
on = False
cur = None
for ln in txt:
    if "Function :" in ln:
        on = key in ln
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m:
        cur = (int(m.group(1), 16), m.group(2).strip())
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m and cur:
        h = int(m.group(1), 16)
        stall, yld, wr, rd, wait = (h >> 41) & 0xF, (h >> 45) & 1, (h >> 46) & 7, (h >> 49) & 7, (h >> 52) & 0x3F
        if lo <= cur[0] < hi:
            print("%06x  st%-2d %s w%s r%s m%02x  %s" % (cur[0], stall, "Y" if yld else "-", "-" if wr == 7 else wr, "-" if rd == 7 else rd, wait, cur[1]))
        cur = None
