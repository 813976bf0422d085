#!/bin/bash
# round 2, closing pass on one GPU: the GPU suite, the secondary configs and an ncu capture of the final box-constrained MPC kernel
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_gpu_tests.log
tail -2 $O/r2_gpu_tests.log
timeout 600 python bench_configs.py > $O/r2_bench_configs.jsonl 2> $O/r2_bench_configs.err
tail -3 $O/r2_bench_configs.err
grep -i "box" $O/r2_bench_configs.jsonl | cut -c1-330
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mpc_track_box -c 1 -f -o $O/r2_mpc_box3 python profiles/box_one.py 1 > $O/r2_ncu_box3.log 2>&1
python profiles/ncu_summarise.py $O/r2_mpc_box3.ncu-rep 10 > $O/r2_mpc_box3_summary.txt
rm -f $O/r2_mpc_box3.ncu-rep
head -16 $O/r2_mpc_box3_summary.txt
