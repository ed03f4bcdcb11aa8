"""One addressing op per mode at a given shape (for an ncu launch list)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import functions as F_, synth
N, M, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 256, 64)
g = torch.Generator().manual_seed(0)
z = torch.randn((1, N, 1, D), generator=g).cuda()
q = A.Quantize_topk(D, M, k=2).cuda().eval()
for mode in ("tensor", "fp32"):
    F_.set_addressing_mode(mode)
    with torch.no_grad():
        for _ in range(2):
            q(z)
    torch.cuda.synchronize()
# module path at the shipped shape
p = synth.memory_params(1)
m = A.enc_quan_dec_res_topk(512, 64, 256, k=2)
m.load_state_dict({"quan." + k: v for k, v in p.items()})
m = m.cuda().eval()
x = synth.features(2, 64).cuda()
for mode in ("tensor", "fp32"):
    F_.set_addressing_mode(mode)
    with torch.no_grad():
        for _ in range(2):
            m(x)
    torch.cuda.synchronize()
