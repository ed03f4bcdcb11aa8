"""Does an SW128 K-major UMMA descriptor accept a start address at any 128-byte row?  (prerequisite of the halo conv)"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ammcnet_aaai2021_b200 import _capi_debug as _capi
from ammcnet_aaai2021_b200.functions import check_pipeline_watchdog
dev = "cuda:0"
rows = 200
g = torch.Generator().manual_seed(0)
a = torch.randint(-4, 5, (rows, 64), generator=g).to(torch.bfloat16).to(dev)
b = torch.randint(-4, 5, (64, 64), generator=g).to(torch.bfloat16).to(dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for off in (0, 1, 2, 3, 7, 8, 9, 15, 33, 34, 66, 72):
    ref = a[off:off + 128].float() @ b.float().t()
    res = []
    for base in sorted({0, off & 7}):
        out = torch.zeros((128, 64), device=dev)
        _capi.call("ammc_debug_desc_probe", P(a), P(b), P(out), rows, off, base, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        res.append((base, bool(torch.equal(out, ref)), int((out != ref).sum())))
    print("row_off", off, [(f"base_offset={b_}", "OK" if ok else f"WRONG({n})") for b_, ok, n in res], flush=True)
check_pipeline_watchdog()
