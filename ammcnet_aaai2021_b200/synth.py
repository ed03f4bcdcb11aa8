"""Deterministic synthetic inputs and random-init weights for the hot path (CPU torch generators).

Datasets, FlowNet2 weights and checkpoints are unavailable offline, so tests, `bench.py` and the golden
fixture generator all draw from here.  Distributions follow SURVEY.md section 8(d):

* path-level features  x_rgb, x_op ~ ReLU(N(0,1))  [b, C, h, w]  (post-ReLU `down3` activations,
  reference Code/models/unet.py:985,992)
* frames  gen, gt ~ U(-1, 1)  [b, 3, 256, 256]  (value range after Normalize(.5,.5),
  reference Code/dataset/two_stream_dataset.py:503-506)
* weights: reference constructor defaults -- Conv2d kaiming-uniform(a=sqrt(5)) for weight and
  U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for bias, bank ~ N(0,1) (unet.py:277), BatchNorm affine = (1, 0);
  `trained_bn=True` instead draws non-trivial BN affine/running statistics, as a trained checkpoint has.

Keys of the returned dicts are the reference `state_dict` names (checkpoint row).
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def conv_default_init(out_c: int, in_c: int, ks: int, g: torch.Generator, bias: bool = True):
    """torch.nn.Conv2d.reset_parameters: both bounds equal 1/sqrt(fan_in)."""
    fan_in = in_c * ks * ks
    bound = 1.0 / math.sqrt(fan_in)
    w = _uniform((out_c, in_c, ks, ks), bound, g)
    b = _uniform((out_c,), bound, g) if bias else None
    return w, b


def memory_params(seed: int, C: int = 512, D: int = 64, M: int = 256, k: int = 2,
                  prefix: str = "") -> Dict[str, torch.Tensor]:
    """Parameters and buffers of one `enc_quan_dec_topk` (reference unet.py:318-324, 268-280)."""
    g = _gen(seed)
    enc_w, enc_b = conv_default_init(D, C, 1, g)
    dec_w, dec_b = conv_default_init(C, k * D, 1, g)
    embed = torch.randn((D, M), generator=g, dtype=torch.float32)
    return {
        prefix + "enc.weight": enc_w, prefix + "enc.bias": enc_b,
        prefix + "quantize.embed": embed,
        prefix + "quantize.cluster_size": torch.zeros(M),
        prefix + "quantize.embed_avg": embed.clone(),
        prefix + "dec.weight": dec_w, prefix + "dec.bias": dec_b,
    }


def amft_params(seed: int, C: int = 512, prefix: str = "", trained_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Parameters and buffers of `bridge(in_c=C)` (reference unet.py:956-960, double_conv 8-16)."""
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {}
    for branch in ("O2F", "F20"):
        for ci, bi in ((0, 1), (3, 4)):
            w, _ = conv_default_init(C, C, 3, g, bias=False)
            p[f"{prefix}{branch}.conv.{ci}.weight"] = w
            if trained_bn:
                p[f"{prefix}{branch}.conv.{bi}.weight"] = 0.5 + torch.rand((C,), generator=g)
                p[f"{prefix}{branch}.conv.{bi}.bias"] = 0.2 * torch.randn((C,), generator=g)
                p[f"{prefix}{branch}.conv.{bi}.running_mean"] = 0.1 * torch.randn((C,), generator=g)
                p[f"{prefix}{branch}.conv.{bi}.running_var"] = 0.05 + 0.2 * torch.rand((C,), generator=g)
            else:
                p[f"{prefix}{branch}.conv.{bi}.weight"] = torch.ones(C)
                p[f"{prefix}{branch}.conv.{bi}.bias"] = torch.zeros(C)
                p[f"{prefix}{branch}.conv.{bi}.running_mean"] = torch.zeros(C)
                p[f"{prefix}{branch}.conv.{bi}.running_var"] = torch.ones(C)
            p[f"{prefix}{branch}.conv.{bi}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return p


def path_params(seed: int, C: int = 512, D: int = 64, M: int = 256, k: int = 2) -> Dict[str, torch.Tensor]:
    """All parameters the hot path reads, under the `twostream` state_dict names."""
    p = {}
    p.update(memory_params(seed * 10 + 1, C, D, M, k, prefix="rgb.vq_down3.quan."))
    p.update(memory_params(seed * 10 + 2, C, D, M, k, prefix="op.vq_down3.quan."))
    p.update(amft_params(seed * 10 + 3, C, prefix="bridge."))
    return p


def double_conv_params(prefix: str, in_c: int, out_c: int, g: torch.Generator,
                       trained_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Parameters and buffers of one `double_conv(in_c, out_c)` (reference unet.py:8-16) under `prefix`.conv.*"""
    p: Dict[str, torch.Tensor] = {}
    for (ci, bi), (cin, cout) in zip(((0, 1), (3, 4)), ((in_c, out_c), (out_c, out_c))):
        w, _ = conv_default_init(cout, cin, 3, g, bias=False)
        p[f"{prefix}.conv.{ci}.weight"] = w
        if trained_bn:
            p[f"{prefix}.conv.{bi}.weight"] = 0.5 + torch.rand((cout,), generator=g)
            p[f"{prefix}.conv.{bi}.bias"] = 0.2 * torch.randn((cout,), generator=g)
            p[f"{prefix}.conv.{bi}.running_mean"] = 0.1 * torch.randn((cout,), generator=g)
            p[f"{prefix}.conv.{bi}.running_var"] = 0.05 + 0.2 * torch.rand((cout,), generator=g)
        else:
            p[f"{prefix}.conv.{bi}.weight"] = torch.ones(cout)
            p[f"{prefix}.conv.{bi}.bias"] = torch.zeros(cout)
            p[f"{prefix}.conv.{bi}.running_mean"] = torch.zeros(cout)
            p[f"{prefix}.conv.{bi}.running_var"] = torch.ones(cout)
        p[f"{prefix}.conv.{bi}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return p


def unet_params(seed: int, in_c: int, out_c: int, D: int = 64, M: int = 256, k: int = 2,
                prefix: str = "") -> Dict[str, torch.Tensor]:
    """Every parameter/buffer of one `UNetMem_v7` stream (reference unet.py:908-922) under the state_dict names."""
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {}
    p.update(double_conv_params(prefix + "inc.conv", in_c, 64, g))
    for name, (ci, co) in (("down1", (64, 128)), ("down2", (128, 256)), ("down3", (256, 512))):
        p.update(double_conv_params(f"{prefix}{name}.mpconv.1", ci, co, g))
    for name, (ci, co) in (("up1", (512, 256)), ("up2", (256, 128)), ("up3", (128, 64))):
        # ConvTranspose2d(ci, ci//2, 2, stride=2): weight [ci, ci//2, 2, 2]; torch's fan_in for it is (ci//2)*4
        bound = 1.0 / math.sqrt((ci // 2) * 4)
        p[f"{prefix}{name}.up.weight"] = _uniform((ci, ci // 2, 2, 2), bound, g)
        p[f"{prefix}{name}.up.bias"] = _uniform((ci // 2,), bound, g)
        p.update(double_conv_params(f"{prefix}{name}.conv", ci, co, g))
    w, b = conv_default_init(out_c, 64, 3, g)
    p[prefix + "outc.weight"], p[prefix + "outc.bias"] = w, b
    p.update(memory_params(seed * 10 + 1, 512, D, M, k, prefix=prefix + "vq_down3.quan."))
    return p


def generator_params(seed: int, in_channel=(12, 6), out_channel=(3, 2), D: int = 64, M: int = 256,
                     k: int = 2) -> Dict[str, torch.Tensor]:
    """state_dict of the shipped `twostream` generator (reference unet.py:967-979, 1241-1249): 222 entries."""
    p = {}
    p.update(unet_params(seed * 7 + 1, in_channel[0], out_channel[0], D, M, k, prefix="rgb."))
    p.update(unet_params(seed * 7 + 2, in_channel[1], out_channel[1], D, M, k, prefix="op."))
    p.update(amft_params(seed * 7 + 3, 512, prefix="bridge."))
    return p


def generator_inputs(seed: int, b: int, h: int = 256, w: int = 256, in_channel=(12, 6)):
    """(rgb_input, op_input) with the loader's value ranges (two_stream_dataset.py:94-95,503-506)."""
    g = _gen(seed)
    rgb = _uniform((b, in_channel[0], h, w), 1.0, g)
    op = torch.randn((b, in_channel[1], h, w), generator=g, dtype=torch.float32) * (4.0 / 256.0)
    op[:, 1::2] = op[:, 0::2] / 256.0
    return rgb, op


# source (h, w) -> target (W, H) cases of the frame / flow preprocessing fixtures: the three datasets' native frame
# sizes plus ragged and degenerate ones (tests/golden/preprocess.npz, oracle/gen_golden.py)
PREPROCESS_CASES = {
    "ped2": ((240, 360), (256, 256)), "avenue": ((360, 640), (256, 256)), "shanghaitech": ((480, 856), (256, 256)),
    "same": ((256, 256), (256, 256)), "up_small": ((37, 53), (64, 48)), "down_odd": ((97, 131), (40, 24)),
    "one_row": ((1, 9), (16, 8)), "one_px": ((1, 1), (8, 8)),
}


def preprocess_inputs(seed: int = 20200525):
    """Yields (name, bgr uint8 [h,w,3], flow float32 [h,w,2], (W, H)) for every PREPROCESS_CASES entry, in order, from
    one numpy Generator stream (decoded frames as TurboJPEG/cv2.imread return them; .flo payloads)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    for name, ((h0, w0), size) in PREPROCESS_CASES.items():
        bgr = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
        flow = (rng.standard_normal((h0, w0, 2)) * 3).astype(np.float32)
        yield name, bgr, flow, size


def features(seed: int, b: int, C: int = 512, h: int = 32, w: int = 32) -> torch.Tensor:
    """ReLU(N(0,1)) bottleneck features, NCHW fp32."""
    return torch.relu(torch.randn((b, C, h, w), generator=_gen(seed), dtype=torch.float32))


def frames(seed: int, b: int, c: int = 3, h: int = 256, w: int = 256, noise: float = 0.1):
    """(gen, gt): gt ~ U(-1,1), gen = clamp(gt + noise*N(0,1)) so PSNR lands in a realistic 20-30 dB range."""
    g = _gen(seed)
    gt = _uniform((b, c, h, w), 1.0, g)
    gen = (gt + noise * torch.randn((b, c, h, w), generator=g)).clamp_(-1.0, 1.0)
    return gen, gt


def objective_inputs(c):
    """Seeded inputs of one objective case (oracle/gen_golden.py gen_objectives and tests/test_gpu_losses.py share them): predicted / target frames and flows,
    FlowNet-style flow pairs, discriminator maps of a real and a generated frame ([b, 1, hd, wd]), a commit-loss scalar."""
    s, b, h, w = c["seed"], c["b"], c["h"], c["w"]
    rgb_out, rgb_tgt = frames(s, b, 3, h, w)
    op_out, op_tgt = frames(s + 100, b, 2, h, w, noise=0.02)
    flow_pred, flow_gt = frames(s + 200, b, 2, h, w, noise=0.05)
    d_gen, d_real = frames(s + 300, b, 1, c["hd"], c["wd"], noise=0.5)
    latent = torch.tensor([0.0371 + 1e-3 * s])
    return dict(flow_pred=flow_pred, flow_gt=flow_gt, rgb_out=rgb_out, rgb_tgt=rgb_tgt, op_out=op_out, op_tgt=op_tgt,
                latent=latent, d_gen=d_gen, d_real=d_real)

# video-length lists of the three datasets (frames per sub-video), recovered from the recorded score
# pickles shipped with the reference (Code/ammcnet_os/model_result_save/*).
PED2_VIDEO_LENGTHS = [180, 180, 150, 180, 150, 180, 180, 180, 120, 150, 180, 180]
