"""BASELINE configs[4]: memory-addressing microbench sweep (queries x items x dim) against the tensor roofline.

For each point: the tcgen05 filter kernel alone (2*N*M*D algorithmic FLOP / kernel time, CUDA events, >= 3 warm-ups,
operands larger than L2 or rotated), the whole exact addressing op (pack + filter + refine) through Quantize_topk, and
the generic fp32 CUDA-core kernel for comparison.  Prints one JSON line per point and a summary table.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ammcnet_aaai2021_b200 as A                       # noqa: E402
from ammcnet_aaai2021_b200 import _capi, functions as F_  # noqa: E402

DEV = torch.device("cuda:0")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"], "measured burst"
    except Exception:
        return 1590.0, "fallback"


def time_fn(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def point(N, M, D, k=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn((N, D), generator=g).to(DEV)
    embed = torch.randn((D, M), generator=g).to(DEV)
    lib = _capi.load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    Mpad = lib.ammc_addr_padded_items(M)
    zp = torch.empty((N, D), dtype=torch.bfloat16, device=DEV)
    bank_t = torch.empty((M, D), device=DEV)
    en2 = torch.empty((M,), device=DEV)
    bank_hi = torch.empty((Mpad, D), dtype=torch.bfloat16, device=DEV)
    en2pad = torch.empty((Mpad,), device=DEV)
    emax = torch.empty((4,), device=DEV)
    cand = torch.empty((N, 24), dtype=torch.int32, device=DEV)
    cnt = torch.empty((N, 2), dtype=torch.int32, device=DEV)
    zn2 = torch.empty((N, 2), device=DEV)
    _capi.call("ammc_addr_pack_queries", P(z), P(zp), P(zn2), N, D, st)
    _capi.call("ammc_addr_pack_bank", P(embed), P(bank_t), P(en2), P(bank_hi), P(en2pad), P(emax), D, M, st)
    flops = 2.0 * N * M * D
    iters = max(3, min(50, int(2e12 / flops)))
    t_filter = time_fn(lambda: _capi.call("ammc_addr_filter", P(zp), P(zn2), P(bank_hi), P(en2pad), P(emax), P(cand),
                                          P(cnt), N, D, M, k, st), iters)
    res_cand = float(cnt[:, 0].clamp(max=24).float().mean())
    decided = float(cnt[:, 1].float().mean())
    q = A.Quantize_topk(D, M, k=k).to(DEV).eval()
    q.embed.copy_(embed)
    z4 = z.view(N // 1024, 32, 32, D) if N % 1024 == 0 else z.view(1, N, 1, D)   # frames of 32x32 queries, as shipped
    res = {"N": N, "M": M, "D": D, "k": k, "filter_ms": t_filter, "filter_tflops": flops / t_filter / 1e9,
           "mean_candidates": res_cand, "rows_decided_by_filter": decided}
    with torch.no_grad():
        F_.set_addressing_mode("tensor")
        res["op_tensor_ms"] = time_fn(lambda: q(z4), max(3, iters // 2))
        res["rescan_rows"] = F_.last_addressing_stats()[0]
        idx_t = q.last_idx.clone()
        if flops <= 3e11:
            F_.set_addressing_mode("fp32")
            res["op_fp32_ms"] = time_fn(lambda: q(z4), 3)
            res["indices_identical"] = bool(torch.equal(idx_t, q.last_idx))
        F_.set_addressing_mode("auto")
    pk, src = peaks()
    res["filter_frac_of_bf16_peak"] = res["filter_tflops"] / pk
    res["op_tensor_tflops"] = flops / res["op_tensor_ms"] / 1e9
    res["peak"] = "%s %.0f TF/s" % (src, pk)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--points", default=None, help="N,M,D;N,M,D;... instead of a built-in grid")
    args = ap.parse_args()
    if args.points:
        grid = [tuple(int(v) for v in pt.split(",")) for pt in args.points.split(";")]
    elif args.quick:
        grid = [(65536, 256, 64), (65536, 2000, 64), (262144, 1024, 128), (65536, 8192, 256), (65536, 2048, 512)]
    else:
        grid = []
        for N in (4096, 16384, 65536, 262144, 1048576):
            for M in (16, 256, 1024, 2048, 8192):
                for D in (64, 128, 256, 512, 1024):
                    if 2.0 * N * M * D <= 6e13 and N * D * 6 <= 8e9:
                        grid.append((N, M, D))
    rows = []
    for (N, M, D) in grid:
        r = point(N, M, D)
        rows.append(r)
        print(json.dumps(r), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)
    print("%9s %6s %5s | %10s %9s %7s | %10s %10s %8s" % ("N", "M", "D", "filter ms", "TFLOP/s", "frac", "op(tc) ms", "op(fp32)", "rescan"))
    for r in rows:
        print("%9d %6d %5d | %10.4f %9.1f %7.3f | %10.4f %10s %8d" % (
            r["N"], r["M"], r["D"], r["filter_ms"], r["filter_tflops"], r["filter_frac_of_bf16_peak"], r["op_tensor_ms"],
            ("%.4f" % r["op_fp32_ms"]) if "op_fp32_ms" in r else "-", r["rescan_rows"]))


if __name__ == "__main__":
    main()
