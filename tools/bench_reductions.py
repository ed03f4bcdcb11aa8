"""North-star kernels (b) and (c) against the HBM roofline: the fused commit-loss / gather backward of the memory module
and the PSNR / score reduction.  Algorithmic bytes are stated per launch; peak = MEASURED_PEAKS.json hbm_gbs (fallback 6650)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
dev = "cuda:0"
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); src = "measured"
except Exception:
    hbm, src = 6650.0, "fallback"


def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# (c) PSNR: 2 * n * 3*256*256 * 4 bytes read
for n in (64, 256, 1024):
    gen, gt = (t.to(dev) for t in synth.frames(5, n))
    ms = timed(lambda: F_.psnr_per_frame(gen, gt))
    by = 2 * gen.numel() * 4
    print(json.dumps({"kernel": "psnr_batch (c)", "frames": n, "ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / hbm,
                      "peak_source": src}), flush=True)
# (c) score reduction on the largest recorded dataset size (shanghaitech T = 40 791 frames in 107 videos): latency-bound
T, V = 40791, 107
lens = np.full(V, T // V); lens[: T - lens.sum()] += 1
off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
img, fea = torch.rand(T, device=dev) * 10 + 20, torch.rand(T, device=dev)
ms = timed(lambda: F_.score_reduce_device(img, fea, off, (0.2, 0.6)))
print(json.dumps({"kernel": "score_reduce (c)", "frames": T, "videos": V, "ms": ms, "note": "2 launches, 326 KB of input: latency-bound"}), flush=True)
# (b) memory-module backward at the shipped shapes, b = 64: reads x, g_out (134 MB each), z, idx; writes gx (134 MB)
b, C, D, M, k = 64, 512, 64, 256, 2
p = synth.memory_params(3, C, D, M, k)
m = A.enc_quan_dec_res_topk(C, D, M, k=k)
m.load_state_dict({"quan." + kk: v for kk, v in p.items()})
m = m.to(dev).train()
x = synth.features(7, b, C, 32, 32).to(dev).requires_grad_(True)
out, diff, q1 = m(x)
g = torch.randn_like(out)


def bwd():
    x.grad = None
    (out * g).sum().backward(retain_graph=True)      # includes torch's mul/sum; the module's backward is one C call


torch.cuda.synchronize()
F_.PROFILE["on"] = False
# time the C call alone through the autograd Function's backward
fn = out.grad_fn
ms_total = timed(bwd, iters=10)
ctx_call = lambda: torch.autograd.grad(out, x, g, retain_graph=True)
ms = timed(ctx_call, iters=10)
by = (3 * x.numel() + 2 * b * 1024 * D) * 4 + b * 1024 * k * 8
print(json.dumps({"kernel": "ammc_mem_bwd (b): commit-loss/gather backward + dec/enc gradients", "frames": b, "ms": ms,
                  "algorithmic_GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / hbm,
                  "note": "bytes = x + g_out read, gx written (fp32 NCHW), z + g_z, top-k indices"}), flush=True)
