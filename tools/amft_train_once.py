"""One AMFT training step (forward + backward) at the shipped shape, for an ncu launch list:  python tools/amft_train_once.py [batch]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = 512
p = synth.amft_params(5, C)
m = A.bridge(in_c=C); m.load_state_dict(p); m = m.cuda().train()
zx = synth.features(1, b, C, 32, 32).cuda().requires_grad_(True)
zy = synth.features(2, b, C, 32, 32).cuda().requires_grad_(True)
g = torch.Generator().manual_seed(3)
rx = torch.randn((b, C, 32, 32), generator=g).cuda(); ry = torch.randn((b, C, 32, 32), generator=g).cuda()
for _ in range(3):
    x, y = m(zx, zy)
    ((x * rx).sum() + (y * ry).sum()).backward()
torch.cuda.synchronize()
