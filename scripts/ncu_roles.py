#!/usr/bin/env python
"""Per-role summary of a k_yuv422* capture: the step loops are found by their barrier; prints executed instructions per
step, FP64 share, samples and the top stall reasons.  usage: ncu -i X.ncu-rep --page source --csv > src.csv; ncu_roles.py src.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and '(' not in h]
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    ins.append((int(r[col['Address']], 16), r[col['Source']].strip(), int(r[col['Instructions Executed']] or 0), int(r[col['# Samples']] or 0),
                {h[6:]: int(r[col[h]] or 0) for h in stall}))
base = ins[0][0]
bars = [a - base for a, s, e, sm, st in ins if s.startswith('BAR')]
tot_e = sum(i[2] for i in ins); tot_s = sum(i[3] for i in ins)
regs = []
for a, s, e, sm, st in ins:
    m = re.match(r"(@!?U?P\d+\s+)?BRA(\.U)?\s+(!?U?P\d+,\s*)?0x([0-9a-f]+)", s)
    if m:
        t = int(m.group(4), 16)
        if t < a: regs.append((t - base, a - base))
best = {}
for t, a in regs:
    nb = [b for b in bars if t <= b <= a]
    if len(nb) == 1 and (nb[0] not in best or (a - t) < (best[nb[0]][1] - best[nb[0]][0])): best[nb[0]] = (t, a)
def opname(s):
    p = s.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
for b, (t, a) in sorted(best.items()):
    sel = [i for i in ins if t <= i[0] - base <= a]
    e = sum(i[2] for i in sel); sm = sum(i[3] for i in sel)
    c = collections.Counter()
    for i in sel:
        for k, v in i[4].items(): c[k] += v
    trips = max(1, max(i[2] for i in ins if i[0] - base == b))
    f64 = sum(i[2] for i in sel if opname(i[1]) in ('DADD', 'DMUL', 'DFMA'))
    xu = sum(i[2] for i in sel if opname(i[1]) in ('I2F', 'F2I'))
    top = ", ".join("%s %d%%" % (k, 100 * v / max(1, sum(c.values()))) for k, v in c.most_common(7))
    print("loop +0x%05x..+0x%05x %5d instr  exec %5.1f%% (%.0f per step: %.0f FP64, %.0f conv) samples %5.1f%%  %s" % (
        t, a, (a - t) // 16 + 1, 100 * e / tot_e, e / trips, f64 / trips, xu / trips, 100 * sm / tot_s, top))
