"""Fused frame losses vs the same losses written with torch ops (what the reference runs), 256x256 RGB frames."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
dev = "cuda:0"
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
for n in (8, 64, 256):
    gen, gt = (t.to(dev) for t in synth.frames(3, n))
    gen.requires_grad_(True)

    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    def ours():
        gen.grad = None
        li, lg = A.frame_losses(gen, gt)
        (li + lg).backward()

    def torch_ops():
        gen.grad = None
        (O.intensity_loss(gen, gt) + O.gradient_loss(gen, gt)).backward()

    a, b = timed(ours), timed(torch_ops)
    by = gen.numel() * 4 * 5            # fwd reads gen+gt, bwd reads gen+gt and writes grad
    print(json.dumps({"frames": n, "fused_fwd_bwd_ms": a, "torch_ops_fwd_bwd_ms": b, "speedup": b / a,
                      "fused_GBps": by / a / 1e6, "frac_of_hbm_peak": by / a / 1e6 / hbm}), flush=True)
