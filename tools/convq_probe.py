"""GPU probe of the precision-2 (fp16 + e4m3) AMFT convolution: error vs float64 and time vs the split-bf16 x3 kernel.

    python tools/convq_probe.py [batch]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from ammcnet_aaai2021_b200 import functions as F_   # noqa: E402

DEV = "cuda:0"


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    C, h, w = 512, 32, 32
    g = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn((b, C, h, w), generator=g))
    wt = (torch.rand((C, C, 3, 3), generator=g) * 2 - 1) / (9 * C) ** 0.5
    scale = (0.5 + torch.rand(C, generator=g)).to(DEV)
    shift = (0.1 * torch.randn(C, generator=g)).to(DEV)
    xd, wd = x.to(DEV), wt.to(DEV)
    xq, wq = F_.pack_nhwc_q(xd), F_.pack_conv_weights_q(wd)
    xp, wp = F_.pack_nhwc(xd), F_.pack_conv_weights(wd)
    out = {}
    y2 = F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=False, precision=2)
    y3 = F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, precision=3)
    torch.cuda.synchronize()
    F_.check_pipeline_watchdog()
    nref = min(b, 2)
    ref = torch.relu(torch.nn.functional.conv2d(x[:nref].double(), wt.double(), padding=1) * scale.cpu().double().view(1, -1, 1, 1)
                     + shift.cpu().double().view(1, -1, 1, 1))
    rms = ref.pow(2).mean().sqrt()
    out["err_p2_max_over_rms"] = float((y2[:nref].cpu().double() - ref).abs().max() / rms)
    out["err_p3_max_over_rms"] = float((y3[:nref].cpu().double() - ref).abs().max() / rms)
    out["err_p2_rms_over_rms"] = float((y2[:nref].cpu().double() - ref).pow(2).mean().sqrt() / rms)
    out["err_p3_rms_over_rms"] = float((y3[:nref].cpu().double() - ref).pow(2).mean().sqrt() / rms)
    res = torch.randn((b, C, h, w), generator=g).to(DEV)
    flops = 2.0 * b * h * w * C * 9 * C
    for name, fn in (("p2_nchw_res", lambda: F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=False, residual=res, precision=2)),
                     ("p2_planes", lambda: F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=True, precision=2)),
                     ("p3_nchw_res", lambda: F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, residual=res, precision=3)),
                     ("p3_planes", lambda: F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=True, precision=3)),
                     ("p1_planes", lambda: F_.conv3x3_bn_relu(xp, wp, scale, shift, to_planes=True, precision=1))):
        ms = timed(fn)
        out[name + "_ms"] = ms
        out[name + "_algorithmic_tflops"] = flops / ms / 1e9
    F_.set_conv_pair_mode(5)              # precision 2 without the dy-halo A tiles (one A tile per tap)
    y2n = F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=False, precision=2)
    out["p2_noreuse_vs_reuse_rel"] = float((y2n - y2).abs().max() / y2.abs().max())
    out["p2_noreuse_planes_ms"] = timed(lambda: F_.conv3x3_bn_relu(xq, wq, scale, shift, to_planes=True, precision=2))
    F_.set_conv_pair_mode(1)
    out["batch"] = b
    print(json.dumps(out))


if __name__ == "__main__":
    main()
