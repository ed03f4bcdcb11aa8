"""Whole twostream generator, eval mode, 256x256 frames: tcgen05 engine vs the cuDNN layers of host_model.

    python tools/generator_bench.py [--batch 16] [--steps 10]

Prints one JSON line per arm (frames/s; CUDA-event timing after warm-up, inputs resident in HBM).
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--size", type=int, default=256)
args = ap.parse_args()
dev = "cuda:0"
m = A.get_twostream()
m.load_state_dict(synth.generator_params(3))
m = m.to(dev).eval()
rgb, op = (t.to(dev) for t in synth.generator_inputs(9, args.batch, args.size, args.size))


def timed(fn, label, extra=None):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = F_.LAUNCHES["count"]
    e0.record()
    for _ in range(args.steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    rec = {"arm": label, "batch": args.batch, "frame": args.size, "ms_per_step": ms, "frames_per_s": args.batch / ms * 1e3,
           "our_launches_per_step": (F_.LAUNCHES["count"] - n0) / args.steps}
    rec.update(extra or {})
    print(json.dumps(rec), flush=True)
    return out


with torch.no_grad():
    eng3 = A.GeneratorEngine(m, precision=3)
    y3 = timed(lambda: eng3(rgb, op), "tcgen05 engine, split-bf16 x3 (fp32 parity)")
    g3 = A.GraphedPath(eng3, [rgb, op])
    timed(lambda: g3.replay(), "tcgen05 engine, split-bf16 x3, CUDA-graph replay")
    m.engine = "cudnn"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yf = timed(lambda: m(rgb, op), "cuDNN fp32 U-Net + this package's path")
    torch.backends.cudnn.allow_tf32 = True
    yt = timed(lambda: m(rgb, op), "cuDNN TF32 U-Net (torch default) + this package's path")
    m.bridge.precision = 1
    eng1 = A.GeneratorEngine(m, precision=1)
    y1 = timed(lambda: eng1(rgb, op), "tcgen05 engine, single bf16 pass (bf16 variant)")
    m.bridge.precision = 3
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(json.dumps({"max_rel_err_vs_cudnn_fp32": {"engine_x3": rel(y3[0], yf[0]), "cudnn_tf32": rel(yt[0], yf[0]),
                                                    "engine_bf16": rel(y1[0], yf[0])}}))
F_.check_pipeline_watchdog()
