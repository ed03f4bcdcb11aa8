"""GPU parity: the U-Net encoder/decoder layers on the tcgen05 conv engine and the whole twostream generator
 against torch fp64, the oracle and the golden fixtures made from the live reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
from conftest import load_golden, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _planes_to_nchw(p):
    return (p[0].double() + p[1].double()).permute(0, 3, 1, 2).cpu()


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("cin,cout,h,w,b", [(64, 64, 3, 256, 2), (64, 128, 2, 160, 1), (128, 64, 5, 136, 1),
                                            (64, 256, 2, 384, 1)])
def test_conv_rows_wider_than_128_pixels(cin, cout, h, w, b):
    """Rows are cut into 128-pixel TMA segments; the last one is zero-filled past the row end and masked."""
    x, wt = torch.relu(_rand((b, cin, h, w), 1)), _rand((cout, cin, 3, 3), 2, (9 * cin) ** -0.5)
    scale, shift = (0.5 + torch.rand(cout)).to(DEV), _rand((cout,), 3, 0.1).to(DEV)
    ref = torch.relu(TF.conv2d(x.double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                     + shift.cpu().double().view(1, -1, 1, 1))
    xp, wp = F_.pack_nhwc(x.to(DEV)), F_.pack_conv_weights(wt.to(DEV))
    out_p = torch.empty((2, b, h, w, cout), dtype=torch.bfloat16, device=DEV)
    out_n = torch.empty((b, cout, h, w), dtype=torch.float32, device=DEV)
    F_.conv_layer(xp, wp, scale, shift, out_planes=out_p, out_nchw=out_n)
    assert_close(out_n.cpu(), ref, 1e-4, "wide.nchw")
    assert_close(_planes_to_nchw(out_p), ref, 1e-4, "wide.planes")
    F_.check_pipeline_watchdog()


@pytest.mark.parametrize("cin,cout,h,w,b,precision", [
    (64, 64, 8, 64, 2, 3),       # weights resident in shared memory (18 tiles)
    (64, 64, 37, 100, 1, 3),     # ragged: partial tiles on both edges
    (64, 64, 16, 256, 1, 1),     # single bf16 pass
    (128, 64, 12, 72, 1, 3),     # two 64-channel blocks, streamed weights
    (64, 128, 9, 128, 2, 3),     # N = 128
    (256, 128, 6, 130, 1, 3),
])
def test_halo_kernel_matches_generic_engine_and_fp64(cin, cout, h, w, b, precision):
    """Wide shallow layers: one halo tile per 64-channel block, nine shifted tap descriptors (csrc/halo_conv.cu)."""
    x, wt = torch.relu(_rand((b, cin, h, w), 21)), _rand((cout, cin, 3, 3), 22, (9 * cin) ** -0.5)
    scale, shift = (0.5 + torch.rand(cout)).to(DEV), _rand((cout,), 23, 0.1).to(DEV)
    xp, wp = F_.pack_nhwc(x.to(DEV)), F_.pack_conv_weights(wt.to(DEV))
    outs = {}
    try:
        for halo in (True, False):
            F_.set_conv_halo_mode(halo)
            out_p = torch.zeros((2, b, h, w, cout + 64), dtype=torch.bfloat16, device=DEV)
            out_n = torch.empty((b, cout, h, w), dtype=torch.float32, device=DEV)
            F_.conv_layer(xp, wp, scale, shift, out_planes=out_p, out_c_off=64, out_nchw=out_n, precision=precision)
            outs[halo] = (out_p, out_n)
    finally:
        F_.set_conv_halo_mode(True)
    F_.check_pipeline_watchdog()
    assert bool((outs[True][0][..., :64] == 0).all())
    if precision == 3:
        ref = torch.relu(TF.conv2d(x.double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                         + shift.cpu().double().view(1, -1, 1, 1))
        assert_close(outs[True][1].cpu(), ref, 1e-4, "halo.nchw")
        assert_close(_planes_to_nchw(outs[True][0][..., 64:]), ref, 1e-4, "halo.planes")
    # same products, different fp32 summation order only
    assert_close(outs[True][1].cpu(), outs[False][1].cpu(), 1e-4, "halo-vs-generic")


def test_conv_channel_windows_make_concat_free():
    """A layer may read a channel window of a wider NHWC buffer and write into one: untouched channels stay untouched."""
    b, h, w = 2, 6, 10
    x, wt = torch.relu(_rand((b, 192, h, w), 4)), _rand((128, 64, 3, 3), 5, (9 * 64) ** -0.5)
    one, zero = torch.ones(128, device=DEV), torch.zeros(128, device=DEV)
    xp = F_.pack_nhwc(x.to(DEV))                                              # 192-channel buffer, layer reads [64, 128)
    out = torch.full((2, b, h, w, 320), 7.0, dtype=torch.bfloat16, device=DEV)  # layer writes [64, 192)
    F_.conv_layer(xp, F_.pack_conv_weights(wt.to(DEV)), one, zero, Cin=64, in_c_off=64, out_planes=out, out_c_off=64)
    ref = torch.relu(TF.conv2d(x[:, 64:128].double(), wt.double(), padding=1))
    assert_close(_planes_to_nchw(out[..., 64:192]), ref, 1e-4, "window.out")
    assert bool((out[..., :64] == 7.0).all()) and bool((out[..., 192:] == 7.0).all())
    # unpack reads the same window back as fp32 NCHW
    assert_close(F_.unpack_nhwc(out, 128, 64).cpu(), ref, 1e-4, "window.unpack")


@pytest.mark.parametrize("cin,cout_valid,act", [(12, 64, 1), (6, 64, 1), (64, 3, 2), (64, 2, 2)])
def test_zero_padded_layers_and_tanh(cin, cout_valid, act):
    """The 12/6-channel network inputs and the 3/2-channel `outc` (+bias, tanh) run as zero-padded 64-channel layers."""
    b, h, w = 2, 8, 24
    x, wt = _rand((b, cin, h, w), 6), _rand((cout_valid, cin, 3, 3), 7, (9 * cin) ** -0.5)
    bias = _rand((cout_valid,), 8, 0.2)
    xp = F_.pack_nhwc_padded(x.to(DEV), 64)
    assert xp.shape == (2, b, h, w, 64) and bool((xp[..., cin:] == 0).all())
    wp = F_.pack_conv_weights_padded(wt.to(DEV), 64, 64)
    shift = torch.zeros(64, device=DEV)
    shift[:cout_valid] = bias.to(DEV)
    out = torch.empty((b, cout_valid, h, w), dtype=torch.float32, device=DEV)
    F_.conv_layer(xp, wp, torch.ones(64, device=DEV), shift, act=act, out_nchw=out, cout_valid=cout_valid)
    ref = TF.conv2d(x.double(), wt.double(), bias.double(), padding=1)
    ref = torch.relu(ref) if act == 1 else torch.tanh(ref)
    assert_close(out.cpu(), ref, 1e-4, "padded")


@pytest.mark.parametrize("cin,h,w,b", [(512, 4, 8, 2), (256, 8, 8, 1), (128, 16, 24, 1), (128, 2, 136, 1)])
def test_transposed_conv_scatter(cin, h, w, b):
    """ConvTranspose2d(cin, cin/2, 2, stride=2) (unet.py:46) as a GEMM with a scattering epilogue, written into the
    upper half of the concat buffer."""
    cout = cin // 2
    x, wt, bias = torch.relu(_rand((b, cin, h, w), 9)), _rand((cin, cout, 2, 2), 10, (4 * cout) ** -0.5), _rand((cout,), 11, 0.1)
    cat = torch.zeros((2, b, 2 * h, 2 * w, 2 * cout), dtype=torch.bfloat16, device=DEV)
    F_.conv_layer(F_.pack_nhwc(x.to(DEV)), F_.pack_convT_weights(wt.to(DEV)), torch.ones(4 * cout, device=DEV),
                  bias.to(DEV).repeat(4).contiguous(), taps=1, act=0, up2x=True, out_planes=cat, out_c_off=cout)
    ref = TF.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2)
    assert_close(_planes_to_nchw(cat[..., cout:]), ref, 1e-4, "convT")
    assert bool((cat[..., :cout] == 0).all())
    F_.check_pipeline_watchdog()


@pytest.mark.parametrize("C,h,w,b,cs,off", [(64, 8, 8, 2, 128, 0), (128, 6, 10, 1, 128, 0), (64, 7, 9, 1, 192, 64)])
def test_maxpool_on_planes(C, h, w, b, cs, off):
    x = _rand((b, cs, h, w), 12)
    xp = F_.pack_nhwc(x.to(DEV))
    out = F_.maxpool2_planes(xp, C, off)
    assert out.shape == (2, b, h // 2, w // 2, C)
    ref = TF.max_pool2d(_planes_to_nchw(xp)[:, off:off + C], 2)
    assert torch.equal(_planes_to_nchw(out), ref)          # max of exactly representable values: bit-exact


def _engine_model(seed, precision=3):
    p = synth.generator_params(seed)
    m = A.get_twostream()
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    return p, m.to(DEV).eval()


@pytest.mark.parametrize("name", ["gen_64", "gen_96x160"])
def test_generator_engine_vs_reference_golden(name):
    """Whole twostream forward (eval) on the tcgen05 engine against outputs of the unmodified reference."""
    c, g = load_golden(name)
    p, m = _engine_model(c["seed"])
    rgb, op = synth.generator_inputs(c["seed"] + 500, c["b"], c["h"], c["w"])
    eng = A.GeneratorEngine(m)
    ry, oy, (rd, od), (rq, oq) = eng(rgb.to(DEV), op.to(DEV))
    assert_close(ry.cpu(), g["rgb_y"], 1e-3, name + ".rgb_y")
    assert_close(oy.cpu(), g["op_y"], 1e-3, name + ".op_y")
    assert_close(rd.cpu(), g["rgb_diff"], 1e-3, name + ".rgb_diff")
    assert_close(od.cpu(), g["op_diff"], 1e-3, name + ".op_diff")
    assert_close(rq.cpu(), g["rgb_q1"], 1e-3, name + ".rgb_q1")
    assert_close(oq.cpu(), g["op_q1"], 1e-3, name + ".op_q1")
    F_.check_pipeline_watchdog()


def test_generator_engine_matches_module_forward_and_oracle():
    """Engine == the module's own forward (cuDNN U-Net + this package's path) == oracle, on 128x128 frames."""
    p, m = _engine_model(41)
    rgb, op = synth.generator_inputs(77, 2, 128, 128)
    eng = A.GeneratorEngine(m)
    ry, oy, (rd, od), _ = eng(rgb.to(DEV), op.to(DEV))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    m.engine = "cudnn"                       # the module's own layers (stock torch.nn on cuDNN, fp32)
    try:
        with torch.no_grad():
            my, mo, (md, mod_), _ = m(rgb.to(DEV), op.to(DEV))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
        m.engine = "tcgen05"
    with torch.no_grad():
        dy = m(rgb.to(DEV), op.to(DEV))[0]   # default route of the module in eval + no_grad = the engine
    assert torch.equal(dy, ry)
    assert_close(ry.cpu(), my.cpu(), 1e-3, "engine-vs-module.rgb")
    assert_close(oy.cpu(), mo.cpu(), 1e-3, "engine-vs-module.op")
    assert_close(rd.cpu(), md.cpu(), 1e-3, "engine-vs-module.diff")
    with torch.no_grad():
        o_ry, o_oy, (o_rd, _), _ = O.twostream_forward(rgb[:1], op[:1], p, 2)
    e1 = eng(rgb[:1].to(DEV), op[:1].to(DEV))
    assert_close(e1[0].cpu(), o_ry, 1e-3, "engine-vs-oracle.rgb")
    assert_close(e1[1].cpu(), o_oy, 1e-3, "engine-vs-oracle.op")
    assert_close(e1[2][0].cpu(), o_rd, 1e-3, "engine-vs-oracle.diff")


def test_generator_engine_refuses_what_it_does_not_cover():
    _, m = _engine_model(43)
    eng = A.GeneratorEngine(m)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        eng(torch.zeros(1, 12, 60, 64, device=DEV), torch.zeros(1, 6, 60, 64, device=DEV))
    m.train()
    with pytest.raises(RuntimeError, match="inference only"):
        eng(torch.zeros(1, 12, 64, 64, device=DEV), torch.zeros(1, 6, 64, 64, device=DEV))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()
        eng(torch.zeros(1, 12, 64, 64), torch.zeros(1, 6, 64, 64))


def test_generator_engine_under_cuda_graph_and_in_video_scorer():
    _, m = _engine_model(45)
    eng = A.GeneratorEngine(m)
    rgb, op = (t.to(DEV) for t in synth.generator_inputs(6, 2, 64, 64))
    eager = [t.clone() for t in (eng(rgb, op)[0], eng(rgb, op)[1])]
    g = A.GraphedPath(eng, [rgb, op])
    rgb2, op2 = (t.to(DEV) for t in synth.generator_inputs(7, 2, 64, 64))
    g(rgb2, op2)
    out = g(rgb, op)
    assert torch.equal(out[0], eager[0]) and torch.equal(out[1], eager[1])
    # the engine stands in for the module wherever a generator is expected
    frames = torch.rand(9, 3, 64, 64, device=DEV) * 2 - 1
    flows = torch.randn(8, 2, 64, 64, device=DEV) * 0.02
    rec_e = A.VideoScorer(eng, batch=4).score_video(frames, flows)
    m.engine = "cudnn"
    rec_m = A.VideoScorer(m, batch=4).score_video(frames, flows)
    for k in ("rgb_img_pred", "rgb_fea_comm", "op_fea_comm"):
        assert_close(rec_e[k], rec_m[k], 1e-3, "scorer." + k)
    F_.check_pipeline_watchdog()


def test_generator_engine_tracks_weight_updates():
    """In-place weight changes (an optimizer step, load_state_dict) must be picked up: packs are keyed by versions."""
    _, m = _engine_model(44)
    eng = A.GeneratorEngine(m)
    rgb, op = synth.generator_inputs(5, 1, 64, 64)
    y0 = eng(rgb.to(DEV), op.to(DEV))[0].clone()
    with torch.no_grad():
        m.rgb.outc.bias.add_(0.25)
    y1 = eng(rgb.to(DEV), op.to(DEV))[0]
    assert float((y1 - y0).abs().max()) > 1e-2
