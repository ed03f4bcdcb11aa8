import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
DEV = "cuda:0"
C, h, w, b = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 5, 7, 3)
p = synth.amft_params(100 + h * w, C)
zx, zy = synth.features(h, b, C, h, w), synth.features(w, b, C, h, w)
p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
m = A.bridge(in_c=C); m.load_state_dict(p); m = m.to(DEV).train()
zxg, zyg = zx.to(DEV).requires_grad_(True), zy.to(DEV).requires_grad_(True)
tx, ty = m(zxg, zyg)
gen = torch.Generator().manual_seed(7)
rx, ry = torch.randn(tx.shape, generator=gen), torch.randn(ty.shape, generator=gen)
((tx * rx.to(DEV)).sum() + (ty * ry.to(DEV)).sum()).backward()
leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if v.is_floating_point() and "running" not in k}
pr = dict(p64); pr.update(leaves)
zx64, zy64 = zx.double().requires_grad_(True), zy.double().requires_grad_(True)
rtx, rty, _ = O.amft_forward(zx64, zy64, pr, training=True)
((rtx * rx.double()).sum() + (rty * ry.double()).sum()).backward()
def rep(name, a, r):
    a, r = a.double().cpu(), r.double()
    e = (a - r).abs(); sc = r.abs().max().item()
    print(f"{name:28s} rel_max_err {e.max().item() / max(sc, 1e-30):.3e}")
    return e, sc
rep("train.x", tx.detach(), rtx.detach()); rep("train.y", ty.detach(), rty.detach())
e, sc = rep("g_zx", zxg.grad, zx64.grad); rep("g_zy", zyg.grad, zy64.grad)
bad = e > 2e-3 * sc
print("g_zx bad frac", bad.float().mean().item(), "per h", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 3)).tolist()],
      "per w", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 2)).tolist()], "per img", bad.float().mean(dim=(1, 2, 3)).tolist())
for name, prm in m.named_parameters():
    rep("g_" + name, prm.grad, leaves[name].grad)
F_.check_pipeline_watchdog()
