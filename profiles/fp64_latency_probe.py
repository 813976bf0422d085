"""DFMA latency / issue-rate probe on the device (acro_bench_fp64_chain).  Run on the GPU box:
    python profiles/fp64_latency_probe.py
Prints cycles per DFMA link for: chains per thread x active lanes per warp x warps per SM."""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import _abi

it = 40000
print("blocks threads chains lanes | cycles per link (per chain) | cycles per warp-DFMA issued")
for blocks, threads in ((148, 32), (148, 128), (148, 512)):
    for chains in (1, 2, 4, 8):
        for lanes in (32, 16, 8):
            out = torch.empty(blocks * threads, dtype=torch.float64, device='cuda')
            cyc = torch.zeros(blocks, dtype=torch.int64, device='cuda')
            for _ in range(2):
                _abi.call("acro_bench_fp64_chain", blocks, threads, it, chains, lanes, C.c_void_p(out.data_ptr()),
                          C.c_void_p(cyc.data_ptr()), None)
            torch.cuda.synchronize()
            per_link = cyc.double().mean().item() / it
            warps_per_smsp = max(1, threads // 128)
            print("%5d %6d %6d %5d | %8.2f | %8.2f" % (blocks, threads, chains, lanes, per_link, per_link / chains / warps_per_smsp))
