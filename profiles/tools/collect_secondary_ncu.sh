#!/bin/bash
# FP64-pipe / DRAM / issue metrics of the kernels behind configs 3, 4, 5 and the large-batch Newton line (run on the GPU box)
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_lqr_track|k_sweep|k_mpc_track_pp|k_mpc_track_shared|k_mpc_track_box|k_newton_ring" --csv --log-file gpurun_out/r1b_secondary_ncu.csv python bench_configs.py --quick > gpurun_out/r1b_secondary_ncu.log 2>&1
tail -2 gpurun_out/r1b_secondary_ncu.log
