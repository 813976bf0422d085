#!/bin/bash
# round 2 evidence pass, part 2: ncu --set full of one kernel (argument: duo | spec); the report is summarised on the box
set -x
O=gpurun_out
mkdir -p $O
if [ "$1" = "duo" ]; then
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_newton_duo -c 1 -f -o $O/r2_newton_duo python bench.py --steps 1 --warmup 0 --iters 4 --no-cpu --no-mpc --quick > $O/r2_ncu_duo.log 2>&1
  python profiles/ncu_to_json.py $O/r2_newton_duo.ncu-rep $O/r2_newton_headline_ncu.json batch=4096 iters=4 n_steps=500 gamma_0=0.1
  python profiles/ncu_summarise.py $O/r2_newton_duo.ncu-rep 30 > $O/r2_newton_duo_ncu_summary.txt
else
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_newton_spec -c 1 -f -o $O/r2_newton_spec python profiles/spec_probe.py --one 1.0 spec 4096 30 > $O/r2_ncu_spec.log 2>&1
  python profiles/ncu_to_json.py $O/r2_newton_spec.ncu-rep $O/r2_newton_spec_ncu.json batch=4096 iters=30 n_steps=500 gamma_0=1.0
  python profiles/ncu_summarise.py $O/r2_newton_spec.ncu-rep 30 > $O/r2_newton_spec_ncu_summary.txt
fi
du -sh $O
