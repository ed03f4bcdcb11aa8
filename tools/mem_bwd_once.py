"""Memory-module backward (ammc_mem_bwd) at the shipped shape, three calls, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/mem_bwd_once.py [batch]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C, D, M, k = 512, 64, 256, 2
p = synth.memory_params(3, C, D, M, k)
m = A.enc_quan_dec_res_topk(C, D, M, k=k)
m.load_state_dict({"quan." + kk: v for kk, v in p.items()})
m = m.cuda().train()
x = synth.features(7, b, C, 32, 32).cuda().requires_grad_(True)
out, diff, q1 = m(x)
g = torch.randn_like(out)
torch.cuda.synchronize()
for _ in range(3):
    torch.autograd.grad(out, x, g, retain_graph=True)
torch.cuda.synchronize()
