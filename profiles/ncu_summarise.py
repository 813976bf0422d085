#!/usr/bin/env python
"""Summarise an ncu report (one kernel): key raw metrics, opcode mix, stall reasons, top stall sites.
Usage: python profiles/ncu_summarise.py gpurun_out/x.ncu-rep [n_top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg', 'lts__t_bytes.sum',
        'l1tex__t_bytes.sum', 'sm__inst_executed_pipe_fp64.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print('==', v[h.index('Kernel Name')])
    for i, n in enumerate(h):
        if n in KEEP:
            print('  %-72s %-10s %s' % (n, u[i], v[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[ie]) for r in data)
samp = sum(int(r[isamp]) for r in data)
print('SASS instructions %d, executed warp-instructions %d, samples %d' % (len(data), tot, samp))
ops, ops_s = collections.Counter(), collections.Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += int(r[ie])
    ops_s[op] += int(r[isamp])
print('opcode mix (share of executed / share of samples):')
for op, c in ops.most_common(18):
    print('  %-10s %6.2f%%  %6.2f%%' % (op, 100 * c / tot, 100 * ops_s[op] / samp))
st = collections.Counter()
for r in data:
    for i in stall_cols:
        st[h[i]] += int(r[i] or 0)
print('stall reasons (% of samples):', {k: round(100 * v / samp, 1) for k, v in st.most_common(8)})
print('top stall sites:')
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]
for i in sorted(idx):
    r = data[i]
    top = max(stall_cols, key=lambda c: int(r[c] or 0))
    print('  %5d %5.2f%% exec=%9s %-18s | %s' % (i, 100 * int(r[isamp]) / samp, r[ie], h[top], r[ia][:80]))
