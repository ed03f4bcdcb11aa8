"""Per-layer time of the generator engine (CUDA events around every conv launch of one eager step)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", type=int, default=3)
args = ap.parse_args()
dev = "cuda:0"
m = A.get_twostream(); m.load_state_dict(synth.generator_params(3)); m = m.to(dev).eval()
m.bridge.precision = args.precision
rgb, op = (t.to(dev) for t in synth.generator_inputs(9, args.batch, 256, 256))
eng = A.GeneratorEngine(m, precision=args.precision)
F_.CONCURRENCY["on"] = False      # one stream: the events around a launch must not see the other network stream's kernels
for _ in range(3): eng(rgb, op)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
F_.PROFILE["on"] = True; F_.PROFILE["events"].clear(); F_.PROFILE["tags"].clear()
e0.record(); eng(rgb, op); e1.record()
torch.cuda.synchronize()
F_.PROFILE["on"] = False
tot = 0.0
print(f"{'layer':44s} {'ms':>8s} {'GFLOP':>8s} {'TF/s(1x)':>9s}")
for (s, e), t in zip(F_.PROFILE["events"], F_.PROFILE["tags"]):
    ms = s.elapsed_time(e); tot += ms
    gf = 2.0 * t["b"] * t["h"] * t["w"] * t["taps"] * t["Cin"] * t["Cout"] / 1e9
    name = f"{'convT' if t['up2x'] else ('3x3' if t['taps']==9 else '1x1')} {t['Cin']}->{t['Cout']} @{t['h']}x{t['w']}"
    print(f"{name:44s} {ms:8.3f} {gf:8.1f} {gf/ms:9.1f}")
print(f"conv launches total {tot:.3f} ms of step {e0.elapsed_time(e1):.3f} ms (eager, event overhead included)")
