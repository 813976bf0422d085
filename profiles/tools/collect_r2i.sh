#!/bin/bash
# round 2: the reworked box-constrained MPC kernel - full GPU suite, probe, configs
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_gpu_tests.log
tail -2 $O/r2_gpu_tests.log
timeout 250 python profiles/box_probe.py > $O/r2_box_probe_after.txt 2>&1
timeout 600 python bench_configs.py > $O/r2_bench_configs.jsonl 2> $O/r2_bench_configs.err
grep -i "box" $O/r2_bench_configs.jsonl | cut -c1-400
tail -3 $O/r2_bench_configs.err
