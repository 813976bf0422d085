#!/bin/bash
# round 2: after the MPC kernel rework (dual sweep, box) - full GPU suite, default bench line, configs, launch list
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_gpu_tests.log
tail -2 $O/r2_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
tail -c 300 $O/r2_bench_1gpu.err
timeout 600 python bench_configs.py > $O/r2_bench_configs.jsonl 2> $O/r2_bench_configs.err
tail -3 $O/r2_bench_configs.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["mpc"]["value"], d["mpc"]["e2e"]["value"], d["mpc"]["e2e"]["blocking"]["value"])
PY
