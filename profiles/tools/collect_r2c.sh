#!/bin/bash
# round 2, third pass: two builds of the speculative kernel (trailer with batched waits behind a call / inlined waits)
set -x
O=gpurun_out
mkdir -p $O
echo "== build B: trailer waits batched behind a call" > $O/r2c_spec_probe.txt
timeout 400 python profiles/spec_probe.py 4096 50 duo,spec1,spec,spec4,spec8 >> $O/r2c_spec_probe.txt 2>&1
echo "== build A: trailer waits inlined (YIELD in the trailer loops)" >> $O/r2c_spec_probe.txt
ACRO_B200_LIB=$PWD/gymnast_optimalcontrol_b200/libacro_b200_A.so timeout 400 python profiles/spec_probe.py 4096 50 spec1,spec,spec4,spec8 >> $O/r2c_spec_probe.txt 2>&1
cat $O/r2c_spec_probe.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > $O/r2c_gpu_tests.log
tail -12 $O/r2c_gpu_tests.log
