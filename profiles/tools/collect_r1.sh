#!/bin/bash
# one measurement pass for profiles/ (run on the GPU box)
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/r1b_gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r1b_bench_1gpu.json 2> $O/r1b_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r1b_bench_reference.json 2> $O/r1b_bench_reference.err
timeout 600 python bench_configs.py > $O/r1b_bench_configs.jsonl 2> $O/r1b_bench_configs.err
timeout 300 python profiles/newton_batch_sweep.py > $O/r1b_batch_sweep.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1b_launches.csv python bench.py --steps 2 --warmup 1 --iters 4 --no-cpu --no-mpc > $O/r1b_ncu_list.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_newton_duo -c 1 -o $O/r1b_newton_duo python bench.py --steps 1 --warmup 0 --iters 4 --no-cpu --no-mpc > $O/r1b_ncu_full.log 2>&1
tail -c 400 $O/r1b_bench_1gpu.json
