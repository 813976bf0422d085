#!/bin/bash
# round 2, second pass: why is the duo path inside k_newton_spec slower than k_newton_duo?  (ncu of 4-iteration launches)
set -x
O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_newton_spec -c 1 -f -o $O/r2b_spec1 python profiles/spec_probe.py --one 0.1 spec1 4096 4 > $O/r2b_ncu_spec1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_newton_spec -c 1 -f -o $O/r2b_spec8 python profiles/spec_probe.py --one 1.0 spec8 4096 6 > $O/r2b_ncu_spec8.log 2>&1
tail -2 $O/r2b_ncu_spec1.log $O/r2b_ncu_spec8.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > $O/r2b_gpu_tests.log
tail -12 $O/r2b_gpu_tests.log
