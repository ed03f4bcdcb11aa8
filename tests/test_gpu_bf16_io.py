"""GPU parity of the bf16 feature-I/O variant (BASELINE configs[2]: "bf16 variants stated separately").

The variant changes the tensors at the module boundary, not the arithmetic: a bf16 input is its own hi plane, sums are formed
in fp32 and rounded once at the store.  So against the fp32 kernels fed the WIDENED input everything must agree exactly --
indices, q1, z, commit loss, the AMFT operand planes -- and the bf16 outputs must be the roundings of the fp32 outputs.
Against the fp32 oracle on the unrounded input the difference is the input rounding (2^-9 relative), stated separately.
"""
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mem(seed, C, D, M, k, res=True):
    p = synth.memory_params(seed, C, D, M, k)
    m = (A.enc_quan_dec_res_topk if res else A.enc_quan_dec_topk)(C, D, M, k=k)
    m.load_state_dict({("quan." if res else "") + kk: v.clone() for kk, v in p.items()})
    return m.to(DEV).eval(), p


@pytest.mark.parametrize("b,C,D,M,k,res,native", [(4, 512, 64, 256, 2, True, True), (3, 256, 64, 100, 3, False, True),
                                                   (2, 512, 64, 2000, 2, True, False), (2, 128, 64, 256, 2, True, False),
                                                   (2, 64, 32, 50, 1, True, False)])
def test_module_bf16_io_equals_fp32_kernels_on_widened_input(b, C, D, M, k, res, native):
    m, _ = _mem(11, C, D, M, k, res)
    x16 = synth.features(12, b, C, 32, 32).to(DEV).to(torch.bfloat16)
    lib = A._capi.load()
    assert bool(lib.ammc_mem_io16_supported(b, 32, 32, C, D, M, k)) == native
    q = m.quan.quantize if res else m.quantize
    with torch.no_grad():
        out16, diff16, q1_16 = m(x16)
        idx16, sse16 = q.last_idx.clone(), q.last_sse_frame.clone()
        planes16 = F_.planes_of(out16, "q") or F_.planes_of(out16, "bf16")
        out32, diff32, q1_32 = m(x16.float())
        idx32, sse32 = q.last_idx.clone(), q.last_sse_frame.clone()
        planes32 = F_.planes_of(out32, "q") or F_.planes_of(out32, "bf16")
    assert out16.dtype == torch.bfloat16 and out16.shape == x16.shape
    assert torch.equal(idx16, idx32)
    assert torch.equal(q1_16, q1_32) and torch.equal(diff16, diff32) and torch.equal(sse16, sse32)
    assert torch.equal(out16, out32.to(torch.bfloat16)), "bf16 out must be the rounding of the fp32 out"
    if planes32 is not None:                 # the AMFT operand planes keep the unrounded fp32 values
        assert planes16 is not None
        a = planes16.buf if isinstance(planes16, F_.QPlanes) else planes16
        c = planes32.buf if isinstance(planes32, F_.QPlanes) else planes32
        n = a.numel() - 16 if isinstance(planes16, F_.QPlanes) else a.numel()
        assert torch.equal(a.view(-1)[:n], c.view(-1)[:n])


def test_bridge_bf16_io_rounds_once():
    """bridge on the bf16 outputs of two memory modules: operand planes from the dec epilogue (unrounded), residuals read
    as bf16, results written as bf16 = the rounding of the fp32 kernel's result for the same planes and residual."""
    b, C = 4, 512
    mr, _ = _mem(21, C, 64, 256, 2)
    mo, _ = _mem(22, C, 64, 256, 2)
    br = A.bridge(C)
    br.load_state_dict(synth.amft_params(23, C))
    br = br.to(DEV).eval()
    xr = synth.features(24, b, C, 32, 32).to(DEV).to(torch.bfloat16)
    xo = synth.features(25, b, C, 32, 32).to(DEV).to(torch.bfloat16)
    with torch.no_grad():
        zx, _, _ = mr(xr)
        zy, _, _ = mo(xo)
        x16, y16 = br(zx, zy)
        px, py = F_.planes_of(zx, "q"), F_.planes_of(zy, "q")
        assert px is not None and py is not None
        x32 = br.O2F.forward_fused(py, zx.float(), 2)
        y32 = br.F20.forward_fused(px, zy.float(), 2)
    assert x16.dtype == torch.bfloat16 and y16.dtype == torch.bfloat16
    assert torch.equal(x16, x32.to(torch.bfloat16)) and torch.equal(y16, y32.to(torch.bfloat16))


def test_bridge_bf16_io_without_attached_planes_and_small_channels():
    """Stand-alone bf16 call (no planes attached: the inputs are widened and packed) and a channel count the CTA-pair kernel
    does not serve (widen / narrow around the fp32 kernels)."""
    for C in (512, 64):
        br = A.bridge(C)
        br.load_state_dict(synth.amft_params(31, C))
        br = br.to(DEV).eval()
        zx = synth.features(32, 2, C, 32, 32).to(DEV).to(torch.bfloat16)
        zy = synth.features(33, 2, C, 32, 32).to(DEV).to(torch.bfloat16)
        with torch.no_grad():
            x16, y16 = br(zx, zy)
            x32, y32 = br(zx.float(), zy.float())
        assert torch.equal(x16, x32.to(torch.bfloat16)) and torch.equal(y16, y32.to(torch.bfloat16))


def test_path_bf16_io_vs_fp32_oracle():
    """The whole path fed bf16 features against the fp32 oracle on the UNROUNDED features: what the input rounding costs
    (index agreement, error on the pixels that kept their indices) -- the figures bench.py reports as bf16_io_variant."""
    b, C = 4, 512
    mr, pr = _mem(41, C, 64, 256, 2)
    x = synth.features(44, b, C, 32, 32)
    with torch.no_grad():
        o16, _, _ = mr(x.to(DEV).to(torch.bfloat16))
        idx16 = mr.quan.quantize.last_idx.cpu()
    ref = O.memory_module_forward(x, pr["enc.weight"], pr["enc.bias"], pr["quantize.embed"], pr["dec.weight"],
                                  pr["dec.bias"], 2, residual=True)
    same = (idx16 == ref["idx_topk"]).all(1)
    assert same.float().mean() > 0.98
    o = o16.float().cpu().permute(0, 2, 3, 1).reshape(-1, C)
    r = ref["out"].permute(0, 2, 3, 1).reshape(-1, C)
    err = (o[same] - r[same]).abs().max() / r.abs().max()
    assert err < 2.0 ** -7, err            # one bf16 rounding of the input feature + one of the output


def test_bf16_io_is_inference_only_and_psnr_accepts_bf16():
    m, _ = _mem(51, 512, 64, 256, 2)
    x16 = synth.features(52, 2, 512, 32, 32).to(DEV).to(torch.bfloat16)
    m.train()
    with pytest.raises(RuntimeError, match="inference variant"):
        m(x16)
    gen, gt = synth.frames(53, 4)
    g16 = gen.to(DEV).to(torch.bfloat16)
    p16 = F_.psnr_per_frame(g16, gt.to(DEV))
    p32 = F_.psnr_per_frame(g16.float(), gt.to(DEV))
    assert torch.equal(p16, p32)
    assert torch.equal(A.narrow_bf16(gen.to(DEV)), gen.to(DEV).to(torch.bfloat16))
    assert torch.equal(A.widen_bf16(g16), g16.float())
