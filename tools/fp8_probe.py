"""Probe for the round-2 conv design: plain fp8 (e4m3) tcgen05 MMAs chained into fp16 MMAs through
scale-input-d.  Exact small-integer inputs, so every mode must reproduce the torch result bit for bit."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ammcnet_aaai2021_b200 import _capi_debug as _capi
from ammcnet_aaai2021_b200.functions import check_pipeline_watchdog
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
a8 = torch.randint(-3, 4, (128, 128), generator=g).float()
b8 = torch.randint(-3, 4, (64, 128), generator=g).float()
a16 = torch.randint(-3, 4, (128, 64), generator=g).float()
b16 = torch.randint(-3, 4, (64, 64), generator=g).float()
a8d, b8d = a8.to(dev).to(torch.float8_e4m3fn).contiguous(), b8.to(dev).to(torch.float8_e4m3fn).contiguous()
a16d, b16d = a16.to(dev).half().contiguous(), b16.to(dev).half().contiguous()
P = lambda t: ctypes.c_void_p(t.data_ptr())
want = {1: a8 @ b8.t(), 2: a16 @ b16.t()}
want[0] = want[1] * 2.0 ** -12 + want[2]
names = {0: "fp8 part * 2^-12 (scale-input-d) + fp16 part", 1: "fp8 (e4m3) MMAs alone", 2: "fp16 MMAs alone"}
for mode in (2, 1, 0):
    out = torch.full((128, 64), float("nan"), device=dev)
    _capi.call("ammc_debug_fp8_probe", P(a8d), P(b8d), P(a16d), P(b16d), P(out), mode,
               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    o = out.cpu()
    print(f"mode {mode} ({names[mode]}): exact = {bool(torch.equal(o, want[mode]))}, max abs diff = {float((o - want[mode]).abs().max()):.3e}", flush=True)
check_pipeline_watchdog()

# issue-to-retire rate of the two kinds on one SM (cycles per MMA; an fp8 MMA covers K = 32, a bf16 one K = 16)
cyc = torch.zeros(1, dtype=torch.int64, device=dev)
for n in (64, 128, 256):
    row = {}
    for fp8 in (0, 1):
        for _ in range(2):
            _capi.call("ammc_debug_mma_rate", P(cyc), fp8, n, 2000, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        row["e4m3 K=32" if fp8 else "bf16 K=16"] = round(int(cyc.item()) / 8000.0, 1)
    print(f"M=128 N={n}: clock64 ticks per MMA {row}", flush=True)
check_pipeline_watchdog()
