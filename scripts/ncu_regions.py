#!/usr/bin/env python
"""Stall samples and executed instructions of a kernel by code region, from `ncu -i X.ncu-rep --page source --csv`.
Regions are the backward-branch loops found in the SASS listing itself (innermost loops, largest first) plus the rest.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python scripts/ncu_regions.py src.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append({"addr": int(r[col["Address"]], 16), "src": r[col["Source"]].strip(), "samples": int(r[col["# Samples"]] or 0),
                "exec": int(r[col["Instructions Executed"]] or 0),
                "stalls": {h[6:]: int(r[col[h]] or 0) for h in hdr if h.startswith("stall_") and "(" not in h}})
base = ins[0]["addr"]
loops = []
for k, i in enumerate(ins):
    m = re.match(r"(@!?U?P\d+\s+)?BRA\s.*0x([0-9a-f]+)", i["src"])
    if m:
        t = int(m.group(2), 16) + 0  # target is an absolute offset within the function listing
        # ncu prints absolute addresses in the Source column for branch targets
        if t < i["addr"] and t >= base:
            loops.append((t, i["addr"]))
# keep innermost loops only (no other loop strictly inside)
inner = [l for l in loops if not any((o[0] >= l[0] and o[1] <= l[1] and o != l) for o in loops)]
inner = sorted(set(inner), key=lambda l: l[0])
tot_s = sum(i["samples"] for i in ins)
tot_e = sum(i["exec"] for i in ins)
print("total: %d instructions, %d samples, %d warp-instructions executed" % (len(ins), tot_s, tot_e))


def report(name, sel):
    s = sum(i["samples"] for i in sel)
    e = sum(i["exec"] for i in sel)
    st = {}
    for i in sel:
        for k, v in i["stalls"].items():
            st[k] = st.get(k, 0) + v
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print("%-34s %5d instr  exec %5.1f%%  samples %5.1f%%  samples/exec %.3f   %s" % (
        name, len(sel), 100.0 * e / max(tot_e, 1), 100.0 * s / max(tot_s, 1), (s / max(tot_s, 1)) / max(e / max(tot_e, 1), 1e-9),
        " ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in top)))


covered = set()
for lo, hi in inner:
    sel = [i for i in ins if lo <= i["addr"] <= hi]
    if len(sel) < 100:
        continue
    report("loop +0x%x..+0x%x" % (lo - base, hi - base), sel)
    covered.update(i["addr"] for i in sel)
report("everything else", [i for i in ins if i["addr"] not in covered])
