#!/usr/bin/env python
"""Key metrics of an `ncu --set full` capture as JSON (what bench.py reads for roofline.traffic / pipe_fp64_pct).
Usage: python profiles/ncu_to_json.py capture.ncu-rep out.json key=value ...   (key=value pairs are stored as "launch":
what the captured launch processed, e.g. batch=4096 iters=4 n_steps=500)"""
import csv
import io
import json
import subprocess
import sys

KEEP = {'gpu__time_duration.sum': 'time_ms', 'dram__bytes_read.sum': 'dram_bytes_read', 'dram__bytes_write.sum': 'dram_bytes_write',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_fp64_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
        'launch__registers_per_thread': 'registers', 'launch__grid_size': 'grid', 'launch__block_size': 'block'}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12, 'ms': 1.0, 'us': 1e-3, 's': 1e3, 'ns': 1e-6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    kernels = []
    for v in rows[2:]:
        k = {'kernel': v[h.index('Kernel Name')]}
        for i, n in enumerate(h):
            if n in KEEP:
                k[KEEP[n]] = float(v[i].replace(',', '')) * (SCALE.get(u[i], 1.0) if KEEP[n] in ('time_ms', 'dram_bytes_read', 'dram_bytes_write') else 1.0)
        kernels.append(k)
    launch = {}
    for kv in sys.argv[3:]:
        a, b = kv.split('=', 1)
        launch[a] = float(b) if b.replace('.', '', 1).isdigit() else b
    json.dump({'source': rep, 'launch': launch, 'kernels': kernels}, open(out, 'w'), indent=1)
    print(json.dumps(kernels[0] if kernels else {}, indent=1))


if __name__ == '__main__':
    main()
