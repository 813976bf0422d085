#!/bin/bash
# bench.py on N GPUs of one box (argument: N), launched the way the driver does
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err
tail -c 400 $O/r2_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["full_returns"]["value"], d["mpc"]["value"], d["mpc"]["e2e"]["value"], d["mpc"]["e2e"]["blocking"]["value"])
print(json.dumps(d.get("strong"))[:900])
PY
