# pipelined solve_mpc_tracking: per-call timings with the default number of hardware queues and with 32
for mc in "" 32; do
  if [ -n "$mc" ]; then export CUDA_DEVICE_MAX_CONNECTIONS=$mc; else unset CUDA_DEVICE_MAX_CONNECTIONS; fi
  timeout 100 python profiles/mpc_e2e_probe.py 2>&1 | grep -E "config|total|^[0-9]+: "
done
