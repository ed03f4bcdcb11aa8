"""One double_conv branch, step by step, against fp64 autograd with retained intermediates."""
import os, sys, torch
import torch.nn.functional as TF
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
DEV = "cuda:0"
C, h, w, b = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 5, 7, 3)
p = synth.amft_params(100 + h * w, C)
zx, zy = synth.features(h, b, C, h, w), synth.features(w, b, C, h, w)
gen = torch.Generator().manual_seed(7)
rx, ry = torch.randn(zx.shape, generator=gen), torch.randn(zx.shape, generator=gen)
def planes(t): return (t[0].float() + t[1].float()).permute(0, 3, 1, 2).double().cpu()
def rep(name, a, r):
    d = (a.double().cpu() - r).abs()
    i = int(d.argmax()); idx = []
    for s in reversed(r.shape): idx.append(i % s); i //= s
    print(f"  {name:10s} rel {float(d.max() / r.abs().max()):.3e} at {tuple(reversed(idx))} n_bad {(d > 1e-3 * r.abs().max()).sum().item()}", flush=True)
for br, u, res, r_out in (("O2F", zy, zx, rx), ("F20", zx, zy, ry)):
    print(br)
    w1, w2 = p[br + ".conv.0.weight"], p[br + ".conv.3.weight"]
    g1, b1, g2, b2 = (p[br + k] for k in (".conv.1.weight", ".conv.1.bias", ".conv.4.weight", ".conv.4.bias"))
    # fp64 reference with retained intermediates
    u64 = u.double().requires_grad_(True)
    y1r = TF.conv2d(u64, w1.double(), padding=1); y1r.retain_grad()
    a1r = torch.relu(TF.batch_norm(y1r, None, None, g1.double(), b1.double(), training=True)); a1r.retain_grad()
    y2r = TF.conv2d(a1r, w2.double(), padding=1); y2r.retain_grad()
    outr = res.double() + torch.relu(TF.batch_norm(y2r, None, None, g2.double(), b2.double(), training=True))
    (outr * r_out.double()).sum().backward()
    # ours
    d = lambda t: t.to(DEV)
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    up = F_.pack_nhwc(d(u))
    y1 = F_.conv3x3_bn_relu(up, F_.pack_conv_weights(d(w1)), one, zero, to_planes=False, precision=3, relu=False)
    sc1, sh1, mu1, is1 = F_.bn_batch_stats(y1, d(g1), d(b1), rm.clone(), rv.clone(), 0.1, 1e-5, True)
    a1p, _, _ = F_.bn_apply(y1, sc1, sh1, relu=True, nhwc=True)
    y2 = F_.conv3x3_bn_relu(a1p, F_.pack_conv_weights(d(w2)), one, zero, to_planes=False, precision=3, relu=False)
    sc2, sh2, mu2, is2 = F_.bn_batch_stats(y2, d(g2), d(b2), rm.clone(), rv.clone(), 0.1, 1e-5, True)
    _, _, out = F_.bn_apply(y2, sc2, sh2, relu=True, f32=True, res=d(res))
    rep("y1", y1, y1r.detach()); rep("a1", planes(a1p), a1r.detach()); rep("y2", y2, y2r.detach()); rep("out", out, outr.detach())
    g_out = d(r_out).contiguous()
    gy2p, gg2, gb2 = F_.bn_backward(g_out, y2, sc2, sh2, mu2, is2, relu=True, training=True)
    rep("gy2", planes(gy2p), y2r.grad)
    gw2 = F_.conv3x3_wgrad(gy2p, a1p, 3)
    g_a1 = F_.conv3x3_bn_relu(gy2p, F_.pack_conv_weights_dgrad(d(w2)), one, zero, to_planes=False, precision=3, relu=False)
    rep("g_a1", g_a1, a1r.grad)
    gy1p, gg1, gb1 = F_.bn_backward(g_a1, y1, sc1, sh1, mu1, is1, relu=True, training=True)
    rep("gy1", planes(gy1p), y1r.grad)
    gw1 = F_.conv3x3_wgrad(gy1p, up, 3)
    g_u = F_.conv3x3_bn_relu(gy1p, F_.pack_conv_weights_dgrad(d(w1)), one, zero, to_planes=False, precision=3, relu=False)
    rep("g_u", g_u, u64.grad)
    # the element where gy1 is worst: was it a ReLU boundary?
    dd = (planes(gy1p) - y1r.grad).abs(); i = int(dd.argmax())
    bn1 = TF.batch_norm(y1r.detach(), None, None, g1.double(), b1.double(), training=True).flatten()[i]
    ours = (y1.double().cpu() * sc1.double().cpu().view(1, -1, 1, 1) + sh1.double().cpu().view(1, -1, 1, 1)).flatten()[i]
    print(f"  pre-ReLU value at worst gy1 element: ref {float(bn1):.3e} ours {float(ours):.3e}")
F_.check_pipeline_watchdog()
