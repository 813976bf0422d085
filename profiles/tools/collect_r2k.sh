#!/bin/bash
# round 2, final one-GPU evidence pass: GPU suite, bench (both arms), secondary configs, ncu launch list of the bench command
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_gpu_tests.log
tail -2 $O/r2_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
tail -c 300 $O/r2_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 600 python bench_configs.py > $O/r2_bench_configs.jsonl 2> $O/r2_bench_configs.err
tail -3 $O/r2_bench_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --iters 4 --no-cpu > $O/r2_ncu_list.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["mpc"]["value"], d["mpc"]["e2e"]["value"], d["mpc"]["e2e"]["blocking"]["value"], d["mpc"]["roofline"]["frac"])
r = json.loads(open("gpurun_out/r2_bench_reference.json").read().strip().splitlines()[-1])
print(r.get("impl"), r.get("value"), r.get("cpu_baseline", {}).get("kind"))
PY
du -sh $O
