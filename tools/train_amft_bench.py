"""Timing of the AMFT block in training mode (forward + backward): this library vs the same block built from stock
torch.nn layers on cuDNN (TF32 allowed / disallowed), same shapes, CUDA events.  Also re-checks gradient agreement."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_

DEV = "cuda:0"


class TorchBridge(torch.nn.Module):      # structure of reference Code/models/unet.py:8-20,956-965 with stock layers
    def __init__(self, c):
        super().__init__()
        def dc():
            return torch.nn.Sequential(torch.nn.Conv2d(c, c, 3, padding=1, bias=False), torch.nn.BatchNorm2d(c), torch.nn.ReLU(True),
                                       torch.nn.Conv2d(c, c, 3, padding=1, bias=False), torch.nn.BatchNorm2d(c), torch.nn.ReLU(True))
        self.O2F, self.F20 = torch.nn.Module(), torch.nn.Module()
        self.O2F.conv, self.F20.conv = dc(), dc()

    def forward(self, zx, zy):
        return zx + self.O2F.conv(zy), zy + self.F20.conv(zx)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


class gm_wrap:
    """a graphed callable with the zero_grad interface `step` uses"""
    def __init__(self, fn, mod):
        self.fn, self.mod = fn, mod

    def __call__(self, *a):
        return self.fn(*a)

    def zero_grad(self, set_to_none=True):
        self.mod.zero_grad(set_to_none=set_to_none)


def main():
    C = 512
    p = synth.amft_params(5, C)
    res = []
    for b in (8, 64):
        zx = synth.features(1, b, C, 32, 32).to(DEV).requires_grad_(True)
        zy = synth.features(2, b, C, 32, 32).to(DEV).requires_grad_(True)
        ours = A.bridge(in_c=C)
        ours.load_state_dict(p)
        ours = ours.to(DEV).train()
        ref = TorchBridge(C)
        ref.load_state_dict(p)
        ref = ref.to(DEV).train()

        g = torch.Generator().manual_seed(3)
        rx = torch.randn((b, C, 32, 32), generator=g).to(DEV)      # random cotangents: a plain sum() loss makes the
        ry = torch.randn((b, C, 32, 32), generator=g).to(DEV)      # BatchNorm backward cancel almost exactly (ill-conditioned)

        def step(m):
            def f():
                for t in (zx, zy):
                    t.grad = None
                m.zero_grad(set_to_none=True)
                x, y = m(zx, zy)
                ((x * rx).sum() + (y * ry).sum()).backward()
            return f
        row = {"batch": b, "ours_ms": timed(step(ours))}
        torch.backends.cudnn.allow_tf32 = True
        row["cudnn_tf32_ms"] = timed(step(ref))
        torch.backends.cudnn.allow_tf32 = False
        row["cudnn_fp32_ms"] = timed(step(ref))
        step(ours)(); g_ours = ours.O2F.conv[0].weight.grad.clone(); gx_ours = zx.grad.clone()
        step(ref)(); g_ref = ref.O2F.conv[0].weight.grad.clone(); gx_ref = zx.grad.clone()
        # ReLU has no derivative at 0: a pre-activation within rounding of 0 may take the other side in the other
        # implementation and changes the gradients by O(1) in its 3x3 neighbourhood, so the max error is dominated by those
        # isolated elements; report the fraction of elements outside 1e-3 next to it, and cuDNN TF32 for calibration
        def cmp(a, r):
            d = (a - r).abs()
            return {"max_rel": float(d.max() / r.abs().max()), "frac_outside_1e-3": float((d > 1e-3 * r.abs().max()).float().mean())}
        row["wgrad_vs_cudnn_fp32"] = cmp(g_ours, g_ref)
        row["gx_vs_cudnn_fp32"] = cmp(gx_ours, gx_ref)
        torch.backends.cudnn.allow_tf32 = True
        step(ref)(); row["cudnn_tf32_gx_vs_cudnn_fp32"] = cmp(zx.grad.clone(), gx_ref)
        row["cudnn_tf32_wgrad_vs_cudnn_fp32"] = cmp(ref.O2F.conv[0].weight.grad.clone(), g_ref)
        torch.backends.cudnn.allow_tf32 = False
        # reduced-precision variant, stated separately: one bf16 tensor-core pass for every convolution of the step
        # (forward, data gradient, weight gradient) -- 8 significant bits per operand against TF32's 10
        ours.precision = 1
        row["ours_single_bf16_pass_ms"] = timed(step(ours))
        step(ours)()
        row["single_bf16_pass_wgrad_vs_cudnn_fp32"] = cmp(ours.O2F.conv[0].weight.grad.clone(), g_ref)
        row["single_bf16_pass_gx_vs_cudnn_fp32"] = cmp(zx.grad.clone(), gx_ref)
        ours.precision = 2
        F_.check_pipeline_watchdog()
        # the same step with forward and backward captured in CUDA graphs (torch.cuda.make_graphed_callables): what is left
        # when the ~60 launches per branch are replayed instead of issued -- matters at the small per-GPU batch of configs[3]
        for name, mod, tf32 in (("ours_graphed_ms", ours, False), ("cudnn_tf32_graphed_ms", ref, True)):
            try:
                torch.backends.cudnn.allow_tf32 = tf32
                gm = torch.cuda.make_graphed_callables(mod, (zx.detach().clone().requires_grad_(True),
                                                             zy.detach().clone().requires_grad_(True)))
                row[name] = timed(step(gm_wrap(gm, mod)))
            except Exception as e:                      # noqa: BLE001  (report, do not hide)
                row[name] = "failed: %s" % (str(e).splitlines()[0][:160],)
        torch.backends.cudnn.allow_tf32 = False
        res.append(row)
        print(json.dumps(row), flush=True)
    json.dump(res, open("gpurun_out/train_amft_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
