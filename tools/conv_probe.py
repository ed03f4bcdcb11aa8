"""GPU diagnostic for the tcgen05 conv engine: prints error structure instead of just pass/fail.
Usage: python tools/conv_probe.py [case ...]   (each case runs in this process; wrap in `timeout`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ammcnet_aaai2021_b200 import functions as F_

DEV = "cuda:0"


def report(name, y, ref):
    y, ref = y.double().cpu(), ref.double().cpu()
    err = (y - ref).abs()
    scale = ref.abs().max().item()
    print(f"[{name}] max_abs_err={err.max().item():.3e} scale={scale:.3e} rel={err.max().item() / scale:.3e} "
          f"finite={bool(torch.isfinite(y).all())} y_absmax={y.abs().max().item():.3e}", flush=True)
    if err.max().item() / scale > 1e-3:
        b, c, h, w = y.shape
        bad = err > 1e-3 * scale
        print("   bad fraction", bad.float().mean().item())
        print("   bad per image  ", bad.float().mean(dim=(1, 2, 3)).tolist())
        pc = bad.float().mean(dim=(0, 2, 3))
        print("   bad per channel (first 16 / nonzero count)", [round(v, 2) for v in pc[:16].tolist()], int((pc > 0).sum()))
        ph = bad.float().mean(dim=(0, 1, 3))
        pw = bad.float().mean(dim=(0, 1, 2))
        print("   bad per row h  ", [round(v, 2) for v in ph.tolist()])
        print("   bad per col w  ", [round(v, 2) for v in pw.tolist()])
        print("   sample y  ", y[0, :4, 0, :4].flatten().tolist())
        print("   sample ref", ref[0, :4, 0, :4].flatten().tolist())
        ratio = (y / ref.clamp_min(1e-9))[ref.abs() > 0.1 * scale]
        if ratio.numel():
            print("   y/ref median", ratio.median().item())


def conv_case(cin, cout, h, w, b, taps, precision):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((b, cin, h, w), generator=g)
    ks = 3 if taps == 9 else 1
    wt = torch.randn((cout, cin, ks, ks), generator=g) / (taps * cin) ** 0.5
    xp = F_.pack_nhwc(x.to(DEV))
    wp = F_.pack_conv_weights(wt.to(DEV))
    one, zero = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    y = F_.conv3x3_bn_relu(xp, wp, one, zero, to_planes=False, precision=precision, relu=False)
    torch.cuda.synchronize()
    if precision == 1:
        xr, wr = x.to(torch.bfloat16).double(), wt.to(torch.bfloat16).double()
    else:
        xr, wr = x.double(), wt.double()
    ref = torch.nn.functional.conv2d(xr, wr, padding=ks // 2)
    report(f"conv{ks}x{ks} cin={cin} cout={cout} {h}x{w} b={b} P={precision}", y, ref)
    yp = F_.conv3x3_bn_relu(xp, wp, one, zero, to_planes=True, precision=precision, relu=False)
    torch.cuda.synchronize()
    rec = (yp[0].float() + yp[1].float()).permute(0, 3, 1, 2)
    report("   -> planes output", rec, ref)


def wgrad_case(cin, cout, h, w, b, precision):
    g = torch.Generator().manual_seed(2)
    x = torch.randn((b, cin, h, w), generator=g)
    gy = torch.randn((b, cout, h, w), generator=g)
    gw = F_.conv3x3_wgrad(F_.pack_nhwc(gy.to(DEV)), F_.pack_nhwc(x.to(DEV)), precision)
    torch.cuda.synchronize()
    F_.check_pipeline_watchdog()
    if precision == 1:
        xr, gr = x.to(torch.bfloat16).double(), gy.to(torch.bfloat16).double()
    else:
        xr, gr = x.double(), gy.double()
    wt = torch.zeros((cout, cin, 3, 3), dtype=torch.float64, requires_grad=True)
    (torch.nn.functional.conv2d(xr, wt, padding=1) * gr).sum().backward()
    ref = wt.grad
    err = (gw.double().cpu() - ref).abs()
    scale = ref.abs().max().item()
    print(f"[wgrad cin={cin} cout={cout} {h}x{w} b={b} P={precision}] max_abs_err={err.max().item():.3e} scale={scale:.3e} "
          f"rel={err.max().item() / scale:.3e}", flush=True)
    if err.max().item() / scale > 1e-3:
        bad = err > 1e-3 * scale
        print("   bad fraction", bad.float().mean().item(), "per tap", bad.float().mean(dim=(0, 1)).flatten().tolist())
        print("   sample gw ", gw[0, :3].flatten().tolist()[:9])
        print("   sample ref", ref[0, :3].flatten().tolist()[:9])


CASES = {
    "w64": lambda: wgrad_case(64, 64, 8, 8, 2, 1),
    "w64b": lambda: wgrad_case(64, 128, 8, 16, 3, 1),
    "w512": lambda: wgrad_case(512, 512, 32, 32, 2, 1),
    "w512p3": lambda: wgrad_case(512, 512, 32, 32, 5, 3),
    "g64": lambda: conv_case(64, 64, 8, 16, 1, 1, 1),
    "g64b": lambda: conv_case(64, 64, 8, 16, 4, 1, 1),
    "g128": lambda: conv_case(128, 128, 8, 16, 2, 1, 1),
    "g256": lambda: conv_case(128, 256, 4, 32, 2, 1, 1),
    "g512": lambda: conv_case(512, 512, 32, 32, 2, 1, 1),
    "c64": lambda: conv_case(64, 64, 8, 16, 2, 9, 1),
    "c64s": lambda: conv_case(64, 64, 8, 8, 3, 9, 1),
    "c512": lambda: conv_case(512, 512, 32, 32, 2, 9, 1),
    "c512p3": lambda: conv_case(512, 512, 32, 32, 2, 9, 3),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    print(torch.cuda.get_device_name(0), flush=True)
    for n in names:
        try:
            CASES[n]()
        except Exception as e:  # keep going: later cases may still tell us something
            print(f"[{n}] EXCEPTION {type(e).__name__}: {e}", flush=True)
            if "CUDA" in str(e) or "cuda" in str(e):
                break
