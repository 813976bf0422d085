#!/bin/bash
# round 2 final pass on one GPU: the whole GPU suite, then the default bench line with the committed library
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
tail -c 600 $O/r2_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["mpc"]["value"], d["mpc"]["e2e"])
PY
du -sh $O
