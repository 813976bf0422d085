#!/bin/bash
set -x
O=gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2e_bench_1gpu.json 2> $O/r2e_bench_1gpu.err
tail -3 $O/r2e_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_1gpu.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync']['ms_per_step'])
print('back',d['backtracking']['value'],d['backtracking']['kernel'],d['backtracking']['ms_per_step'])
print('mpc',d['mpc']['value'],d['mpc']['e2e'])
print('long',{k:(v['value'],v['kernel']) for k,v in d['long_horizon'].items() if k.startswith('batch')})
print('strong',{k:v.get('value') for k,v in d['strong'].items() if isinstance(v,dict)})
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['kind'])
PY
