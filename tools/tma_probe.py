"""GPU: what does a 5-D TMA box with a narrow inner dimension look like in shared memory (128B swizzle)?"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ammcnet_aaai2021_b200 import _capi_debug as _capi, functions as F_

def probe(W, H, C, B, box, coords, tag):
    lib = _capi.load()
    n = 2 * B * C * H * W
    # element value = its linear index within the plane-0 tensor (mod 2^15), exactly representable ids via int16 view
    ids = (torch.arange(n, dtype=torch.int32) % 32000).to(torch.int16).cuda()
    dims = (ctypes.c_int64 * 5)(W, H, C, B, 2)
    strides = (ctypes.c_int64 * 4)(W * 2, H * W * 2, C * H * W * 2, B * C * H * W * 2)
    bx = (ctypes.c_int * 5)(*box)
    co = (ctypes.c_int * 5)(*coords)
    nbytes = int(np.prod(box)) * 2
    out = torch.full((nbytes,), 0x77, dtype=torch.uint8, device="cuda")
    rc = lib.ammc_debug_tma_probe(ctypes.c_void_p(ids.data_ptr()), dims, strides, bx, co, ctypes.c_void_p(out.data_ptr()),
                                  nbytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    print(tag, "rc", rc, lib.ammc_last_error() if rc else "")
    try:
        torch.cuda.synchronize()
        F_.check_pipeline_watchdog()
    except Exception as e:
        print(tag, "FAULT:", str(e).splitlines()[0])
        return
    got = out.view(torch.int16).cpu().numpy()
    # expected dense packing (w fastest, then h, then c) with the 128B swizzle on 16-byte chunks
    exp = np.zeros(int(np.prod(box)), np.int16)
    it = 0
    for c in range(box[2]):
        for h in range(box[1]):
            for w in range(box[0]):
                ww, hh, cc = coords[0] + w, coords[1] + h, coords[2] + c
                v = 0
                if 0 <= ww < W and 0 <= hh < H and 0 <= cc < C:
                    v = (((coords[3] * C + cc) * H + hh) * W + ww) % 32000
                byte = it * 2
                chunk = (byte >> 4)
                sw = chunk ^ ((byte >> 7) & 7)
                exp[(sw << 4 | (byte & 15)) >> 1] = v
                it += 1
    ok = np.array_equal(got, exp)
    print(tag, "dense+swizzle layout matches:", ok)
    if not ok:
        print("  got[:48]", got[:48].tolist())
        print("  exp[:48]", exp[:48].tolist())

print(torch.cuda.get_device_name(0))
cases = [
    (64, 8, 8, 2, (64, 8, 2, 1, 1), (0, 0, 0, 0, 0), "known-good conv-like inner=128B"),          # 0
    (64, 8, 8, 2, (64, 8, 2, 1, 1), (0, -1, 1, 0, 0), "conv-like shifted"),                        # 1
    (32, 8, 64, 2, (32, 2, 128, 1, 1), (0, 0, 0, 0, 0), "inner=64B noshift box C 128 > 64"),       # 2
    (32, 8, 128, 2, (32, 2, 128, 1, 1), (0, 0, 0, 0, 0), "inner=64B noshift C=128"),               # 3
    (32, 8, 128, 2, (32, 2, 128, 1, 1), (0, 1, 0, 0, 0), "inner=64B dy+1"),                        # 4
    (32, 8, 128, 2, (32, 2, 128, 1, 1), (-1, 0, 0, 0, 0), "inner=64B dx-1"),                       # 5
    (32, 8, 128, 2, (32, 2, 128, 1, 1), (8, 0, 0, 0, 0), "inner=64B dx+8 (16B aligned)"),          # 6
    (8, 8, 128, 2, (8, 8, 128, 1, 1), (0, 0, 0, 0, 0), "inner=16B noshift"),                       # 7
    (8, 8, 64, 2, (8, 8, 128, 1, 1), (0, 0, 0, 0, 0), "A 8x8 noshift"),                                 # 8
    (8, 8, 64, 2, (8, 8, 128, 1, 1), (-1, -1, 0, 0, 0), "8x8 dx-1 dy-1"),
    (8, 8, 64, 2, (8, 8, 128, 1, 1), (0, 1, 0, 0, 0), "8x8 dy+1"),
    (8, 8, 64, 2, (8, 8, 128, 1, 1), (1, 1, 0, 1, 0), "8x8 dx+1 dy+1 img1"),
    (32, 32, 64, 1, (32, 2, 128, 1, 1), (-1, 1, 0, 0, 0), "32x32 dx-1 dy+1"),
    (32, 32, 64, 1, (32, 2, 128, 1, 1), (1, 31, 0, 0, 0), "32x32 dx+1 last rows"),
]
only = sys.argv[1:]
for i, c in enumerate(cases):
    if only and str(i) not in only:
        continue
    probe(*c)
