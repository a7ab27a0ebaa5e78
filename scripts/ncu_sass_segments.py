#!/usr/bin/env python
"""Condensed view of an `ncu --page source --csv` (SASS) export: stall samples summed between
marker instructions (TMEM loads/stores, barriers, MMAs).  usage: ncu_sass_segments.py export.csv [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
hdr = next(i for i, r in enumerate(rows) if '# Samples' in r)
H = rows[hdr]
iS, iI, iSrc = H.index('# Samples'), H.index('Instructions Executed'), H.index('Source')
ins = []
for r in rows[hdr + 1:]:
    try:
        ins.append((r[iSrc].strip(), int(r[iS]), int(r[iI])))
    except (ValueError, IndexError):
        pass
print(len(ins), 'instructions,', sum(s for _, s, _ in ins), 'samples')
markers = ('LDTM', 'STTM', 'SYNCS', 'BAR.SYNC', 'UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'EXIT', 'WARPSYNC', 'FENCE')
seg = [0, 0, 0]
out = []
for src, s, i in ins:
    op = src.split()[1] if src.startswith('@') else src.split()[0]
    if any(op.startswith(m) for m in markers):
        if seg[2]:
            out.append(('  ...%d instr' % seg[2], seg[0], seg[1] // seg[2]))
        out.append((src[:70], s, i))
        seg = [0, 0, 0]
    else:
        seg[0] += s
        seg[1] += i
        seg[2] += 1
if seg[2]:
    out.append(('  ...%d instr' % seg[2], seg[0], seg[1] // seg[2]))
for o in out:
    if o[1] >= thr:
        print('%7d %11d  %s' % (o[1], o[2], o[0]))
