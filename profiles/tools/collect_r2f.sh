#!/bin/bash
# round 2 evidence pass (one GPU), part 1: tests, bench (both arms), secondary configs, probes, ncu launch list
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_gpu_tests.log
tail -3 $O/r2_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 600 python bench_configs.py > $O/r2_bench_configs.jsonl 2> $O/r2_bench_configs.err
tail -2 $O/r2_bench_configs.err
timeout 300 python profiles/spec_probe.py 4096 50 > $O/r2_spec_probe.txt 2>&1
timeout 300 python profiles/e2e_probe.py > $O/r2_e2e_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --iters 4 --no-cpu > $O/r2_ncu_list.log 2>&1
ACRO_B200_LIB=$PWD/gymnast_optimalcontrol_b200/libacro_b200_T.so timeout 120 python profiles/spec_timing.py 0.1 1 > $O/r2_spec_timing_g01.txt 2>&1
ACRO_B200_LIB=$PWD/gymnast_optimalcontrol_b200/libacro_b200_T.so timeout 120 python profiles/spec_timing.py 1.0 0 > $O/r2_spec_timing_g1.txt 2>&1
head -30 $O/r2_spec_timing_g01.txt
du -sh $O
