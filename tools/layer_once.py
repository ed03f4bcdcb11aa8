"""Run one conv layer a few times (for ncu):  python tools/layer_once.py Cin Cout h w b [halo=1] [precision=3]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ammcnet_aaai2021_b200 import functions as F_
cin, cout, h, w, b = (int(a) for a in sys.argv[1:6])
halo = int(sys.argv[6]) if len(sys.argv) > 6 else 1
prec = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = "cuda:0"
x = torch.relu(torch.randn(b, cin, h, w, device=dev))
wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
xp, wp = F_.pack_nhwc(x), F_.pack_conv_weights(wt)
one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
out = torch.empty((2, b, h, w, cout), dtype=torch.bfloat16, device=dev)
F_.set_conv_halo_mode(bool(halo))
for _ in range(3):
    F_.conv_layer(xp, wp, one, zero, out_planes=out, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    F_.conv_layer(xp, wp, one, zero, out_planes=out, precision=prec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{cin}->{cout} @{h}x{w} b={b} halo={halo} prec={prec}: {ms:.3f} ms, {2*b*h*w*9*cin*cout/ms/1e9:.1f} TF/s (1x)")
F_.check_pipeline_watchdog()
