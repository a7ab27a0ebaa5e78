#!/usr/bin/env python
"""Rank CUDA source lines of an `ncu --page source --print-source cuda,sass --csv` export by
warp-stall samples.  usage: ncu_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if '# Samples' in r)
H = rows[hdr]
ix = {n: H.index(n) for n in ['# Samples', 'Instructions Executed', 'L1 Wavefronts Shared Excessive']}
stalls = [(i, n) for i, n in enumerate(H) if n.startswith('stall_') and 'Not Issued' not in n]
agg = {}
fname = ''
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    if len(r) < len(H) or not r[0].isdigit():
        continue
    try:
        s = int(r[ix['# Samples']])
    except ValueError:
        continue
    key = (fname, int(r[0]), r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0, 0, {}])
    a[0] += s
    a[1] += int(r[ix['Instructions Executed']] or 0)
    a[2] += int(r[ix['L1 Wavefronts Shared Excessive']] or 0)
    for i, n in stalls:
        try:
            a[3][n] = a[3].get(n, 0) + int(r[i])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
print('total samples', tot, ' total warp-instructions', sum(a[1] for a in agg.values()))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print('%5.1f%% %9d inst %9d exc  %s:%d  %s   [%s]' % (100.0 * a[0] / tot, a[1], a[2], key[0], key[1], key[2][:70],
                                                      ', '.join('%s %d' % (n[6:], v) for n, v in st)))
