#!/bin/bash
# round 2, first measurement pass (run on the GPU box): tests, bench (both arms)
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2a_gpu_tests.log
tail -3 $O/r2a_gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2a_bench_1gpu.json 2> $O/r2a_bench_1gpu.err
tail -c 1500 $O/r2a_bench_1gpu.json; tail -5 $O/r2a_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2a_bench_reference.json 2> $O/r2a_bench_reference.err
tail -c 800 $O/r2a_bench_reference.json; tail -3 $O/r2a_bench_reference.err
