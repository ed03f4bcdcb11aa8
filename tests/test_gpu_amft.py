"""GPU parity: AMFT block (tcgen05 implicit-GEMM convolutions) against the golden fixtures and the oracle."""
import numpy as np
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
from conftest import load_golden, assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bridge(c, p, precision):
    m = A.bridge(in_c=c["C"], precision=precision)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    return m.to(DEV).eval()


def _inputs(c):
    p = synth.amft_params(c["seed"], c["C"])
    zx = synth.features(c["seed"] + 1000, c["b"], c["C"], c["h"], c["w"])
    zy = synth.features(c["seed"] + 2000, c["b"], c["C"], c["h"], c["w"])
    return p, zx, zy


def test_pack_nhwc_roundtrip():
    x = synth.features(5, 3, 96, 6, 10).to(DEV)
    xp = F_.pack_nhwc(x)
    assert xp.shape == (2, 3, 6, 10, 96) and xp.dtype == torch.bfloat16
    rec = (xp[0].float() + xp[1].float()).permute(0, 3, 1, 2)
    assert (rec - x).abs().max() <= 2.0 ** -16 * x.abs().max()
    assert torch.equal(xp[0], x.permute(0, 2, 3, 1).to(torch.bfloat16))


@pytest.mark.parametrize("cin,cout,h,w,b", [(64, 64, 8, 16, 2), (128, 256, 4, 32, 1), (64, 128, 8, 8, 3), (512, 512, 32, 32, 2)])
def test_conv1x1_engine(cin, cout, h, w, b):
    """The tensor-core engine without the conv halo: a plain GEMM, hi-plane only, must match fp64 on bf16 inputs."""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn((b, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, 1, 1), generator=g) / cin ** 0.5
    xp = F_.pack_nhwc(x.to(DEV))
    wp = F_.pack_conv_weights(wt.to(DEV))
    one, zero = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    y = F_.conv3x3_bn_relu(xp, wp, one, zero, to_planes=False, precision=1, relu=False)
    xb, wb = x.to(torch.bfloat16).double(), wt.to(torch.bfloat16).double()
    ref = torch.nn.functional.conv2d(xb, wb)
    assert rel_err(y.cpu(), ref) < 1e-5
    y3 = F_.conv3x3_bn_relu(xp, wp, one, zero, to_planes=False, precision=3, relu=False)
    assert rel_err(y3.cpu(), torch.nn.functional.conv2d(x.double(), wt.double())) < 5e-5


@pytest.mark.parametrize("cin,cout,h,w,b", [(64, 64, 8, 8, 2), (64, 64, 8, 8, 3), (128, 64, 4, 32, 2), (64, 256, 16, 16, 1), (512, 512, 32, 32, 1)])
def test_conv3x3_engine(cin, cout, h, w, b):
    g = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = torch.randn((b, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, 3, 3), generator=g) / (9 * cin) ** 0.5
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV)
    shift = torch.randn(cout, generator=g).to(DEV)
    res = torch.randn((b, cout, h, w), generator=g)
    xp = F_.pack_nhwc(x.to(DEV))
    wp = F_.pack_conv_weights(wt.to(DEV))
    ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                     + shift.cpu().double().view(1, -1, 1, 1))
    y = F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, residual=res.to(DEV), precision=3)
    assert_close(y.cpu(), ref + res.double(), 1e-3, "conv3x3.nchw")
    assert rel_err(y.cpu(), ref + res.double()) < 5e-5
    yp = F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=True, precision=3)
    rec = (yp[0].float() + yp[1].float()).permute(0, 3, 1, 2)
    assert rel_err(rec.cpu(), ref) < 5e-5
    y1 = F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, precision=1)
    assert rel_err(y1.cpu(), ref) < 2e-2          # bf16 variant, stated separately


def test_cta_pair_kernel_matches_single_cta_kernel():
    """cta_group::2 (M=256 over two SMs) and the single-CTA kernel accumulate in the same order -> identical bits."""
    g = torch.Generator().manual_seed(9)
    for (cin, cout, h, w, b) in [(64, 256, 16, 16, 3), (512, 512, 32, 32, 3), (128, 256, 8, 8, 5)]:
        x = torch.randn((b, cin, h, w), generator=g)
        wt = torch.randn((cout, cin, 3, 3), generator=g) / (9 * cin) ** 0.5
        xp, wp = F_.pack_nhwc(x.to(DEV)), F_.pack_conv_weights(wt.to(DEV))
        scale, shift = torch.rand(cout, generator=g).to(DEV) + 0.5, torch.randn(cout, generator=g).to(DEV)
        outs = []
        try:
            for pair in (3, False, True):
                F_.set_conv_pair_mode(pair)
                outs.append((F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, precision=3),
                             F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=True, precision=1)))
        finally:
            F_.set_conv_pair_mode(True)
        F_.check_pipeline_watchdog()
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        # default kernel (hi/lo planes fused per K block): same products, different fp32 summation order
        assert rel_err(outs[2][0].cpu(), outs[1][0].cpu()) < 2e-5 and torch.equal(outs[2][1], outs[1][1])
        ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                         + shift.cpu().double().view(1, -1, 1, 1))
        assert rel_err(outs[0][0].cpu(), ref) < 5e-5


@pytest.mark.parametrize("name,precision", [("amft_c64", 3), ("amft_c512", 3), ("amft_c512", 2), ("amft_c64", 2)])
def test_bridge_vs_golden(name, precision):
    """Eval forward against the live-reference fixture in both fp32-parity modes: 3 = split-bf16 x3, 2 = fp16 + e4m3 cross
    terms (C=64 is outside the q kernel's shapes: the module must fall back to precision 3 there, not fail)."""
    c, g = load_golden(name)
    p, zx, zy = _inputs(c)
    m = _bridge(c, p, precision=precision)
    assert m.eval_precision(c["C"]) == (precision if c["C"] % 256 == 0 else 3)
    with torch.no_grad():
        x, y = m(zx.to(DEV), zy.to(DEV))
    F_.check_pipeline_watchdog()
    assert_close(x.cpu(), g["x"], 1e-3, name + ".x")
    assert_close(y.cpu(), g["y"], 1e-3, name + ".y")
    e = max(rel_err(x.cpu(), g["x"]), rel_err(y.cpu(), g["y"]))
    print(name, "precision", precision, "rel err", e)
    assert e < 3e-4                                  # both parity modes sit well inside the 1e-3 bar
    m1 = _bridge(c, p, precision=1)
    with torch.no_grad():
        x1, y1 = m1(zx.to(DEV), zy.to(DEV))
    e = max(rel_err(x1.cpu(), g["x"]), rel_err(y1.cpu(), g["y"]))
    print(name, "bf16 variant rel err", e)
    assert e < 3e-2


@pytest.mark.parametrize("cin,cout,h,w,b,xs", [(128, 256, 12, 12, 3, 1.0), (512, 512, 32, 32, 2, 1.0), (256, 256, 5, 7, 2, 1e-4),
                                               (128, 512, 3, 40, 1, 3e4)])
def test_q_conv_engine_vs_fp64(cin, cout, h, w, b, xs):
    """precision 2 on its own: q operands (fp16 + two e4m3 planes, power-of-two scales found on the device) against fp64,
    at input magnitudes from 1e-4 to 3e4 (fp16 alone would under/overflow without the scales)."""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn((b, cin, h, w), generator=g) * xs
    x[0, 0, 0, 0] = 40.0 * xs                        # an outlier sets the scale; the bulk sits 5 binades below it
    wt = torch.randn((cout, cin, 3, 3), generator=g) / (9 * cin) ** 0.5
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV) / xs
    shift = torch.randn(cout, generator=g).to(DEV)
    res = torch.randn((b, cout, h, w), generator=g)
    xq = F_.pack_nhwc_q(x.to(DEV))
    assert (xq.dequantize().cpu() - x).abs().max() <= 2.0 ** -14 * x.abs().max()     # h16 + l8/16 carries ~15 bits
    s16 = float(xq.scale())
    assert 2 ** 14 <= float(x.abs().max()) * s16 < 2 ** 15 and s16 == 2.0 ** round(np.log2(s16))
    wq = F_.pack_conv_weights_q(wt.to(DEV))
    ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                     + shift.cpu().double().view(1, -1, 1, 1))
    y = F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=False, residual=res.to(DEV), precision=2)
    F_.check_pipeline_watchdog()
    assert_close(y.cpu(), ref + res.double(), 1e-3, "q.conv3x3.nchw")
    e = rel_err(y.cpu() - res, ref)
    print("q conv", (cin, cout, h, w, b, xs), "rel err", e)
    assert e < 1.5e-4
    yq = F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=True, precision=2)       # q planes out: scale from the L1 bound
    assert isinstance(yq, F_.QPlanes)
    assert float(ref.abs().max()) * float(yq.scale()) < 2 ** 15
    assert rel_err(yq.dequantize().cpu(), ref) < 2e-4


def test_bridge_full_size_linearity_and_oracle_sample():
    """b=8 at the shipped 512x32x32 shape: oracle on 1 frame (frames are independent in eval mode) + batch-split invariance
    (precision 3: the q path's scales are per-tensor, so its bits depend on the batch's max -- covered below)."""
    C = 512
    p = synth.amft_params(91, C)
    zx, zy = synth.features(92, 8, C, 32, 32), synth.features(93, 8, C, 32, 32)
    ox, oy, _ = O.amft_forward(zx[3:4], zy[3:4], p)
    for prec in (3, 2):
        m = A.bridge(in_c=C, precision=prec)
        m.load_state_dict(p)
        m = m.to(DEV).eval()
        with torch.no_grad():
            x, y = m(zx.to(DEV), zy.to(DEV))
            xa, ya = m(zx[3:5].to(DEV), zy[3:5].to(DEV))
        if prec == 3:
            assert torch.equal(xa, x[3:5]) and torch.equal(ya, y[3:5])
        else:
            assert rel_err(xa.cpu(), x[3:5].cpu()) < 1e-4 and rel_err(ya.cpu(), y[3:5].cpu()) < 1e-4
        assert_close(x[3:4].cpu(), ox, 1e-3, "bridge.full.x")
        assert_close(y[3:4].cpu(), oy, 1e-3, "bridge.full.y")


@pytest.mark.parametrize("name", ["amft_c64", "amft_c512"])
def test_bridge_training_forward_backward_vs_reference_autograd(name):
    """Train-mode forward (batch-statistic BN, running-stat update) and every gradient against the reference's autograd."""
    c, g = load_golden(name)
    p, zx, zy = _inputs(c)
    m = A.bridge(in_c=c["C"], precision=3)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    m = m.to(DEV).train()
    zxg, zyg = zx.to(DEV).requires_grad_(True), zy.to(DEV).requires_grad_(True)
    tx, ty = m(zxg, zyg)
    assert_close(tx.detach().cpu(), g["train_x"], 1e-3, name + ".train_x")
    assert_close(ty.detach().cpu(), g["train_y"], 1e-3, name + ".train_y")
    sd = m.state_dict()
    for k in g:
        if k.startswith("stat_"):
            assert_close(sd[k[5:]].cpu(), g[k], 1e-3, name + "." + k)
    assert int(sd["O2F.conv.1.num_batches_tracked"]) == 1
    gen = torch.Generator().manual_seed(c["seed"] + 3000)
    rx, ry = torch.randn(tx.shape, generator=gen).to(DEV), torch.randn(ty.shape, generator=gen).to(DEV)
    ((tx * rx).sum() + (ty * ry).sum()).backward()
    assert_close(zxg.grad.cpu(), g["g_zx"], 1e-3, name + ".g_zx")
    assert_close(zyg.grad.cpu(), g["g_zy"], 1e-3, name + ".g_zy")
    import digest
    n_checked = 0
    for i, (pn, pv) in enumerate(m.named_parameters()):
        if "g_" + pn in g:
            assert_close(pv.grad.cpu(), g["g_" + pn], 1e-3, name + ".g_" + pn)
            n_checked += 1
        elif "gd_rows_" + pn in g:        # shipped C=512: reduced forms of the [512,512,3,3] gradients (oracle/digest.py)
            for dk, dv in digest.weight_grad_digest(pv.grad, c["seed"] + 4000 + i).items():
                assert_close(dv, g["gd_%s_%s" % (dk, pn)], 1e-3, "%s.gd_%s_%s" % (name, dk, pn))
            n_checked += 1
    assert n_checked == 12


def test_bridge_eval_mode_with_grad_matches_fused_path():
    """eval() + autograd (frozen BN) must give the same forward as the fused no-grad path and finite gradients."""
    c, g = load_golden("amft_c64")
    p, zx, zy = _inputs(c)
    m = _bridge(c, p, precision=3)
    with torch.no_grad():
        x0, y0 = m(zx.to(DEV), zy.to(DEV))
    zxg = zx.to(DEV).requires_grad_(True)
    x1, y1 = m(zxg, zy.to(DEV))
    assert_close(x1.detach().cpu(), x0.cpu(), 1e-4, "eval-grad.x")
    assert_close(y1.detach().cpu(), y0.cpu(), 1e-4, "eval-grad.y")
    (x1.sum() + y1.sum()).backward()
    ref_m = torch.nn.Sequential()          # fp64 torch reference of the same frozen-BN block, for the input gradient
    import ammc_oracle as O2
    zr = zx.double().requires_grad_(True)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
    ox, oy, _ = O2.amft_forward(zr, zy.double(), p64)
    (ox.sum() + oy.sum()).backward()
    assert_close(zxg.grad.cpu(), zr.grad, 1e-3, "eval-grad.g_zx")


def test_bridge_rejects_unsupported_shapes():
    m96 = A.bridge(in_c=96).to(DEV)
    with torch.no_grad(), pytest.raises(RuntimeError, match="multiples of 64"):
        m96.eval()(torch.zeros(1, 96, 8, 8, device=DEV), torch.zeros(1, 96, 8, 8, device=DEV))
    # rows wider than 128 pixels: inference works (128-pixel segments), the weight-gradient kernel does not cover them
    m = A.bridge(in_c=64).to(DEV)
    z = torch.randn(1, 64, 4, 130, device=DEV)
    with torch.no_grad():
        x, _ = m.eval()(z, z)
    ref, _, _ = O.amft_forward(z.cpu().double(), z.cpu().double(), {k: (v.double().cpu() if v.is_floating_point() else v.cpu())
                                                                     for k, v in m.state_dict().items()})
    assert_close(x.cpu(), ref, 1e-3, "bridge.w130")
    with pytest.raises(RuntimeError, match="128 pixels wide"):
        out, _ = m.train()(z.requires_grad_(True), z)
        out.sum().backward()


RELU_BAND = 5e-5      # |fp64 pre-activation| below this (BatchNorm output, unit scale): the ReLU side is not decidable in fp32


def _relu_masks_of_branch(u, w1, g1, b1, w2, g2, b2):
    """The ReLU activity masks of the CUDA training path (same kernels, same inputs => same bits)."""
    C = u.shape[1]
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    d = lambda t: t.to(DEV)
    if F_.q_conv_supported(C, C):       # the training path of bridge(precision=2) runs these layers on q operands
        y1 = F_.conv3x3_bn_relu(F_.pack_nhwc_q(d(u)), F_.pack_conv_weights_q(d(w1)), one, zero, to_planes=False, precision=2,
                                relu=False)
        sc1, sh1, _, _, ws1 = F_.bn_batch_stats_q(y1, d(g1), d(b1), rm.clone(), rv.clone(), 0.1, 1e-5)
        _, a1q = F_.bn_apply_q(y1, sc1, sh1, ws1, relu=True, nhwc=False)
        _, _, a1 = F_.bn_apply(y1, sc1, sh1, relu=True, f32=True)
        y2 = F_.conv3x3_bn_relu(a1q, F_.pack_conv_weights_q(d(w2)), one, zero, to_planes=False, precision=2, relu=False)
    else:
        y1 = F_.conv3x3_bn_relu(F_.pack_nhwc(d(u)), F_.pack_conv_weights(d(w1)), one, zero, to_planes=False, precision=3,
                                relu=False)
        sc1, sh1, _, _ = F_.bn_batch_stats(y1, d(g1), d(b1), rm.clone(), rv.clone(), 0.1, 1e-5, True)
        a1p, _, a1 = F_.bn_apply(y1, sc1, sh1, relu=True, nhwc=True, f32=True)
        y2 = F_.conv3x3_bn_relu(a1p, F_.pack_conv_weights(d(w2)), one, zero, to_planes=False, precision=3, relu=False)
    sc2, sh2, _, _ = F_.bn_batch_stats(y2, d(g2), d(b2), rm.clone(), rv.clone(), 0.1, 1e-5, True)
    _, _, a2 = F_.bn_apply(y2, sc2, sh2, relu=True, f32=True)
    return (a1 > 0).cpu(), (a2 > 0).cpu()


def _double_conv_ref_banded(u, p, prefix, cuda_masks, stats):
    """oracle.double_conv_forward in training mode (unet.py:8-20) in float64.  ReLU has no derivative at 0: where the
    fp64 pre-activation lies within RELU_BAND of 0 an fp32 implementation may legitimately land on either side, and the
    gradients then differ by O(1) at that element (and, through the weight gradient, at a whole row).  Those elements --
    and ONLY those -- take the side the CUDA path took; everywhere else the fp64 sign decides and the CUDA mask must
    agree with it.  `stats` collects (elements inside the band, elements, mask disagreements outside the band)."""
    for (ci, bi), cm in zip(((0, 1), (3, 4)), cuda_masks):
        u = torch.nn.functional.conv2d(u, p[f"{prefix}.conv.{ci}.weight"], None, padding=1)
        u = torch.nn.functional.batch_norm(u, None, None, p[f"{prefix}.conv.{bi}.weight"], p[f"{prefix}.conv.{bi}.bias"],
                                           training=True, eps=1e-5)
        band = u.detach().abs() < RELU_BAND
        m64 = u.detach() > 0
        stats[0] += int(band.sum()); stats[1] += band.numel(); stats[2] += int(((m64 != cm) & ~band).sum())
        u = u * torch.where(band, cm, m64).to(u.dtype)
    return u


@pytest.mark.parametrize("C,h,w,b", [(64, 5, 7, 3), (128, 12, 20, 2), (64, 9, 100, 1), (256, 28, 28, 2), (64, 3, 128, 2),
                                     (64, 1, 1, 5)])
def test_bridge_arbitrary_feature_map_sizes(C, h, w, b):
    """Any width <= 128 / any height (partial TMA boxes, masked rows): eval forward, train forward and all gradients
    against the oracle differentiated by torch autograd in float64."""
    p = synth.amft_params(100 + h * w, C)
    zx, zy = synth.features(h, b, C, h, w), synth.features(w, b, C, h, w)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
    m = A.bridge(in_c=C)
    m.load_state_dict(p)
    m = m.to(DEV).eval()
    with torch.no_grad():
        x, y = m(zx.to(DEV), zy.to(DEV))
    ox, oy, _ = O.amft_forward(zx.double(), zy.double(), p64)
    assert_close(x.cpu(), ox, 1e-3, "odd.eval.x")
    assert_close(y.cpu(), oy, 1e-3, "odd.eval.y")
    if b * h * w < 2:
        return                                           # BatchNorm needs more than one value per channel to train
    # training mode + gradients
    m.train()
    zxg, zyg = zx.to(DEV).requires_grad_(True), zy.to(DEV).requires_grad_(True)
    tx, ty = m(zxg, zyg)
    gen = torch.Generator().manual_seed(7)
    rx, ry = torch.randn(tx.shape, generator=gen), torch.randn(ty.shape, generator=gen)
    ((tx * rx.to(DEV)).sum() + (ty * ry.to(DEV)).sum()).backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if v.is_floating_point() and "running" not in k}
    pr = dict(p64)
    pr.update(leaves)
    zx64, zy64 = zx.double().requires_grad_(True), zy.double().requires_grad_(True)
    # plain oracle for the forward values; for the gradients the ReLU sides are taken from the CUDA path (see above)
    otx, oty, _ = O.amft_forward(zx.double(), zy.double(), p64, training=True)
    assert_close(tx.detach().cpu(), otx, 1e-3, "odd.train.x")
    assert_close(ty.detach().cpu(), oty, 1e-3, "odd.train.y")
    br = lambda name: [p[name + k] for k in (".conv.0.weight", ".conv.1.weight", ".conv.1.bias", ".conv.3.weight",
                                              ".conv.4.weight", ".conv.4.bias")]
    masks_o = _relu_masks_of_branch(zy, *br("O2F"))
    masks_f = _relu_masks_of_branch(zx, *br("F20"))
    st = [0, 0, 0]
    rtx = zx64 + _double_conv_ref_banded(zy64, pr, "O2F", masks_o, st)
    rty = zy64 + _double_conv_ref_banded(zx64, pr, "F20", masks_f, st)
    assert st[2] == 0, "ReLU sides differ from float64 outside the undecidable band at %d elements" % st[2]
    assert st[0] <= max(2, 1e-4 * st[1]), "band holds %d of %d elements" % (st[0], st[1])
    assert_close(rtx.detach(), otx, 1e-4, "odd.banded-reference.x")     # the band only moves values within 5e-5 of 0
    ((rtx * rx.double()).sum() + (rty * ry.double()).sum()).backward()
    assert_close(zxg.grad.cpu(), zx64.grad, 1e-3, "odd.g_zx")
    assert_close(zyg.grad.cpu(), zy64.grad, 1e-3, "odd.g_zy")
    for name, prm in m.named_parameters():
        assert_close(prm.grad.cpu(), leaves[name].grad, 1e-3, "odd.g_" + name)
    F_.check_pipeline_watchdog()


def test_bridge_training_gradients_at_shipped_shape_vs_float64():
    """C=512, 32x32 (the cta_group::2 pair kernel + split-K tcgen05 weight gradient): train forward and ALL gradients --
    both inputs, the four [512,512,3,3] conv weights, the eight BatchNorm vectors -- against float64 autograd of the oracle,
    ReLU sides fixed by the fp64-defined band above."""
    C, h, w, b = 512, 32, 32, 2
    p = synth.amft_params(77, C)
    zx, zy = synth.features(78, b, C, h, w), synth.features(79, b, C, h, w)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
    m = A.bridge(in_c=C)
    m.load_state_dict(p)
    m = m.to(DEV).train()
    zxg, zyg = zx.to(DEV).requires_grad_(True), zy.to(DEV).requires_grad_(True)
    tx, ty = m(zxg, zyg)
    gen = torch.Generator().manual_seed(7)
    rx, ry = torch.randn(tx.shape, generator=gen), torch.randn(ty.shape, generator=gen)
    ((tx * rx.to(DEV)).sum() + (ty * ry.to(DEV)).sum()).backward()
    F_.check_pipeline_watchdog()
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if v.is_floating_point() and "running" not in k}
    pr = dict(p64)
    pr.update(leaves)
    zx64, zy64 = zx.double().requires_grad_(True), zy.double().requires_grad_(True)
    br = lambda name: [p[name + k] for k in (".conv.0.weight", ".conv.1.weight", ".conv.1.bias", ".conv.3.weight",
                                              ".conv.4.weight", ".conv.4.bias")]
    st = [0, 0, 0]
    rtx = zx64 + _double_conv_ref_banded(zy64, pr, "O2F", _relu_masks_of_branch(zy, *br("O2F")), st)
    rty = zy64 + _double_conv_ref_banded(zx64, pr, "F20", _relu_masks_of_branch(zx, *br("F20")), st)
    assert st[2] == 0, "ReLU sides differ from float64 outside the undecidable band at %d elements" % st[2]
    assert st[0] <= 1e-4 * st[1], "band holds %d of %d elements" % (st[0], st[1])
    assert_close(tx.detach().cpu(), rtx.detach(), 1e-3, "c512.train.x")
    assert_close(ty.detach().cpu(), rty.detach(), 1e-3, "c512.train.y")
    ((rtx * rx.double()).sum() + (rty * ry.double()).sum()).backward()
    assert_close(zxg.grad.cpu(), zx64.grad, 1e-3, "c512.g_zx")
    assert_close(zyg.grad.cpu(), zy64.grad, 1e-3, "c512.g_zy")
    n = 0
    for name, prm in m.named_parameters():
        assert_close(prm.grad.cpu(), leaves[name].grad, 1e-3, "c512.g_" + name)
        n += 1
    assert n == 12


def test_staged_batchnorm_statistics_equal_single_call():
    """Global-batch BatchNorm plumbing (ammc_bn_*_staged): with a stand-in all-reduce that doubles the per-channel sums and
    reports two ranks -- the statistics of the batch seen twice -- forward, gradients and running means must equal the
    single-call per-rank path (the unbiased running variance differs by the n/(n-1) factor of the doubled count)."""
    c, g = load_golden("amft_c64")
    p, zx, zy = _inputs(c)
    res = {}
    for mode in ("local", "staged"):
        m = A.bridge(in_c=c["C"], precision=3)
        m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
        m = m.to(DEV).train()
        F_.BN_SYNC["allreduce"] = (lambda sums: (sums.mul_(2.0), 2)[1]) if mode == "staged" else None
        try:
            zxg, zyg = zx.to(DEV).requires_grad_(True), zy.to(DEV).requires_grad_(True)
            tx, ty = m(zxg, zyg)
            (tx.sum() + (ty * ty).sum()).backward()
        finally:
            F_.BN_SYNC["allreduce"] = None
        res[mode] = (tx.detach(), ty.detach(), zxg.grad, zyg.grad, [q.grad for q in m.parameters()],
                     m.O2F.conv[1].running_mean.clone(), m.O2F.conv[1].running_var.clone())
    a, b = res["local"], res["staged"]
    for i in range(4):
        assert torch.equal(a[i], b[i]), i
    for ga, gb in zip(a[4], b[4]):
        assert torch.equal(ga, gb)
    assert torch.equal(a[5], b[5])
    n = c["b"] * c["h"] * c["w"]
    assert_close(b[6] - 0.9 * p["O2F.conv.1.running_var"].to(DEV), (a[6] - 0.9 * p["O2F.conv.1.running_var"].to(DEV)) *
                 ((2 * n) / (2 * n - 1.0)) / (n / (n - 1.0)), 1e-5, "running_var with the doubled count")
