#!/usr/bin/env python
"""Attribute ncu source-page samples to the loops of a kernel.  Usage: ncu_loops.py source_page.csv [detail_loop_index]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
addr = [int(r[0], 16) for r in data]
tot = sum(int(r[isamp]) for r in data)
loops = []
for i, r in enumerate(data):
    if 'BRA' in r[ia]:
        m = re.search(r'0x([0-9a-f]+)', r[ia])
        if m and int(m.group(1), 16) < addr[i] and int(m.group(1), 16) in addr:
            j = addr.index(int(m.group(1), 16))
            if i - j > 150 and int(r[ie]) > 1000: loops.append((j, i, int(r[ie])))
for n, (j, i, e) in enumerate(loops):
    ss = sum(int(data[k][isamp]) for k in range(j, i + 1)); ee = sum(int(data[k][ie]) for k in range(j, i + 1))
    fp = sum(int(data[k][ie]) for k in range(j, i + 1) if re.match(r'\s*(@!?U?P\d+\s+)?D(FMA|MUL|ADD)', data[k][ia]))
    print('loop %d: SASS %d-%d, trips/warp-sum %d, %.1f%% of samples, %.0f instr/trip, %.0f FP64/trip' % (n, j, i, e, 100 * ss / tot, ee / e, fp / e))
if len(sys.argv) > 2:
    j, i, e = loops[int(sys.argv[2])]
    thr = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    for k in range(j, i + 1):
        r = data[k]; s = int(r[isamp])
        top = max(stall_cols, key=lambda c: int(r[c] or 0))
        if s > thr: print('%5d %5d %-16s %s' % (k, s, h[top].replace('stall_', ''), r[ia][:95]))
