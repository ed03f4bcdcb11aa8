"""Step through AmftBranchFn.backward by hand and compare every intermediate with fp64 torch."""
import os, sys, torch
import torch.nn.functional as TF
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
DEV = "cuda:0"
C, h, w, b = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 5, 7, 3)
g = torch.Generator().manual_seed(1)
gy = torch.randn((b, C, h, w), generator=g)
a1 = torch.relu(torch.randn((b, C, h, w), generator=g))
wt = torch.randn((C, C, 3, 3), generator=g) / (9 * C) ** 0.5
one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
gyp = F_.pack_nhwc(gy.to(DEV)); a1p = F_.pack_nhwc(a1.to(DEV))
ref_gx = TF.conv_transpose2d(gy.double(), wt.double(), padding=1)          # data gradient of conv2d(padding=1)
wd = wt.double().clone().requires_grad_(True)
(TF.conv2d(a1.double(), wd, padding=1) * gy.double()).sum().backward()
def rel(a, r): return float((a.double().cpu() - r).abs().max() / r.abs().max())
for trial in range(3):
    for order in ("wgrad-then-dgrad", "dgrad-only", "dgrad-then-wgrad"):
        if order == "wgrad-then-dgrad":
            gw = F_.conv3x3_wgrad(gyp, a1p, 3)
            gx = F_.conv3x3_bn_relu(gyp, F_.pack_conv_weights_dgrad(wt.to(DEV)), one, zero, to_planes=False, precision=3, relu=False)
        elif order == "dgrad-only":
            gx = F_.conv3x3_bn_relu(gyp, F_.pack_conv_weights_dgrad(wt.to(DEV)), one, zero, to_planes=False, precision=3, relu=False)
            gw = None
        else:
            gx = F_.conv3x3_bn_relu(gyp, F_.pack_conv_weights_dgrad(wt.to(DEV)), one, zero, to_planes=False, precision=3, relu=False)
            gw = F_.conv3x3_wgrad(gyp, a1p, 3)
        torch.cuda.synchronize()
        print(trial, order, "dgrad rel err", f"{rel(gx, ref_gx):.3e}", "wgrad rel err", "-" if gw is None else f"{rel(gw, wd.grad):.3e}", flush=True)
F_.check_pipeline_watchdog()
