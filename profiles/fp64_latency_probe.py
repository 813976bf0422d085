import ctypes as C, torch, sys
sys.path.insert(0,'.')
from gymnast_optimalcontrol_b200 import _abi
for blocks,threads in ((1,32),(148,32),(148,128),(148,256),(148,512)):
    out=torch.empty(blocks*threads,dtype=torch.float64,device='cuda'); cyc=torch.zeros(blocks,dtype=torch.int64,device='cuda')
    it=80000
    for _ in range(2):
        _abi.call("acro_bench_fp64_chain",blocks,threads,it,C.c_void_p(out.data_ptr()),C.c_void_p(cyc.data_ptr()),None)
    torch.cuda.synchronize()
    print(blocks,threads,'cycles per dependent DFMA: %.2f'%(cyc.double().mean().item()/it))
