"""torch.autograd.Function wrappers around the C ABI (raw device pointers + the current CUDA stream).

PyTorch is plumbing here: it owns device memory (caching allocator), streams and autograd bookkeeping; all
arithmetic of the path runs in libammc_b200.so.  Inputs must be CUDA fp32 tensors -- anything else raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _capi


# --------------------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(*tensors, names=()):
    for i, t in enumerate(tensors):
        if t is None:
            continue
        nm = names[i] if i < len(names) else "tensor %d" % i
        if not t.is_cuda:
            raise RuntimeError("ammc_b200: %s must live on a CUDA device (there is no CPU path)" % nm)
        if t.dtype != torch.float32:
            raise RuntimeError("ammc_b200: %s must be float32, got %s" % (nm, t.dtype))


_checked_devices = set()


def _check_device(dev: torch.device):
    if dev.index in _checked_devices:
        return
    with torch.cuda.device(dev):
        if not _capi.load().ammc_device_supported():
            raise RuntimeError("ammc_b200: device %s is not sm_100 (B200); this library has no other target" % dev)
    _checked_devices.add(dev.index)


_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}
_retired_workspaces = []          # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream), owned by PyTorch's caching allocator.  A buffer that is outgrown
    is retired, never freed: CUDA graphs captured earlier (GraphedPath, VideoScorer(graph=True)) replay launches that
    carry its raw address."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired_workspaces.append(ws)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_side_streams: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}
CONCURRENCY = {"on": True}       # False: `concurrently` runs its pieces one after the other on the current stream (per-kernel timing)


def concurrently(*fns):
    """Run independent pieces of the path on separate CUDA streams and join them on the current stream.

    The appearance and the motion memory module do not depend on each other until the AMFT block (reference
    unet.py:985-994), and neither does the PSNR of the previous prediction; their many short kernels (bank preparation,
    per-frame reductions) and the tails of the large ones overlap when issued on different streams.  fns[0] runs on the
    current stream, the others on per-device side streams; works eagerly and under CUDA-graph capture (fork/join).
    Returns the list of results."""
    if not CONCURRENCY["on"]:
        return [fn() for fn in fns]
    cur = torch.cuda.current_stream()
    dev = cur.device
    results = [None] * len(fns)
    sides = []
    for i in range(1, len(fns)):
        key = (dev.index, i)
        st = _side_streams.get(key)
        if st is None:
            st = torch.cuda.Stream(device=dev)
            _side_streams[key] = st
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            results[i] = fns[i]()
        sides.append(st)
    results[0] = fns[0]()
    for st in sides:
        cur.wait_stream(st)
    if not torch.cuda.is_current_stream_capturing():     # a captured graph owns its memory pool: nothing is ever reused
        def mark(obj):                                    # tensors produced on a side stream are consumed on this one
            if isinstance(obj, (torch.Tensor, QPlanes)):
                obj.record_stream(cur)
                planes = getattr(obj, "_ammc_planes", None)
                if planes is not None:
                    planes[0].record_stream(cur)           # bf16 planes tensor or QPlanes
            elif isinstance(obj, (tuple, list)):
                for o in obj:
                    mark(o)
            elif isinstance(obj, dict):
                for o in obj.values():
                    mark(o)
        for r in results[1:]:
            mark(r)
    return results


# launch counter: bench.py reports how many of OUR kernels ran in the timed region
LAUNCHES = {"count": 0}


def _count(n):
    LAUNCHES["count"] += n


# optional CUDA-event bracketing of the dominant kernel (bench.py's roofline measurement, live in the timed region)
PROFILE = {"on": False, "events": [], "tags": []}


# --------------------------------------------------------------------------------------------------
# memory module
# --------------------------------------------------------------------------------------------------
def set_dec_mode(mode: str = "auto"):
    """'auto' | 'fp32' (table gather on CUDA cores) | 'tensor' (split-bf16 x3 GEMM on tcgen05); process-wide."""
    _capi.call("ammc_set_dec_mode", {"auto": 0, "fp32": 1, "tensor": 2}[mode])


def set_enc_mode(mode: str = "auto"):
    """'auto' | 'fp32' (FFMA GEMM on CUDA cores) | 'tensor' (split-bf16 x3 on tcgen05, NCHW converted on the fly)."""
    _capi.call("ammc_set_enc_mode", {"auto": 0, "fp32": 1, "tensor": 2}[mode])


class QPlanes:
    """q-format operand buffer of an NHWC activation [b,h,w,C] (include/ammc_b200.h, precision 2): fp16 plane, two e4m3
    planes and the power-of-two scale in one uint8 tensor."""
    __slots__ = ("buf", "shape")

    def __init__(self, buf: torch.Tensor, shape):
        self.buf, self.shape = buf, tuple(shape)          # shape = (b, h, w, C)

    @staticmethod
    def empty(b, h, w, C, device):
        n = b * h * w * C
        return QPlanes(torch.empty((4 * n + 16,), dtype=torch.uint8, device=device), (b, h, w, C))

    @property
    def device(self):
        return self.buf.device

    def data_ptr(self):
        return self.buf.data_ptr()

    def record_stream(self, st):
        self.buf.record_stream(st)

    def scale(self) -> torch.Tensor:
        n = self.buf.numel() - 16
        return self.buf[n:n + 4].view(torch.float32)

    def dequantize(self) -> torch.Tensor:
        """fp32 NCHW value of the fp16 plane + residual plane (tests / debugging)."""
        b, h, w, C = self.shape
        n = b * h * w * C
        h16 = self.buf[:2 * n].view(torch.float16).float()
        l8 = self.buf[3 * n:4 * n].view(torch.float8_e4m3fn).float()
        return ((h16 + l8 / 16.0) / self.scale()).view(b, h, w, C).permute(0, 3, 1, 2).contiguous()


def attach_planes(t: torch.Tensor, planes):
    """Remember the NHWC operand planes of `t` (bf16 hi/lo tensor or QPlanes, produced for free by the dec epilogue) so the
    AMFT block can skip its pack kernel.  The tensor's version counter is recorded: any in-place change of `t`
    invalidates the planes."""
    t._ammc_planes = (planes, t._version)


def planes_of(t: torch.Tensor, fmt: str = "bf16"):
    """The attached operand planes of `t` in format `fmt` ('bf16' | 'q'), or None."""
    rec = getattr(t, "_ammc_planes", None)
    if rec is None:
        return None
    planes, version = rec
    b, C, h, w = t.shape
    if version != t._version or planes.device != t.device:
        return None
    if isinstance(planes, QPlanes):
        return planes if fmt == "q" and planes.shape == (b, h, w, C) else None
    if fmt != "bf16" or tuple(planes.shape) != (2, b, h, w, C):
        return None
    return planes


def set_front_mode(on: bool = True):
    """True (default): eval forwards at the shipped shapes run enc + addressing + exact refine as one persistent kernel
    (csrc/mem_front.cu); False: the staged kernels.  For A/B measurements; process-wide."""
    _capi.call("ammc_set_front_mode", int(bool(on)))


def mem_prepare(enc_w, embed, dec_w, dec_b, b: int, h: int, w: int, k: int) -> torch.Tensor:
    """Parameter-only quantities of the memory module (packed weights, bank rows / norms / bf16 copy, q-scale bound) in one
    buffer; valid until one of the four tensors changes.  `enc_w` [D,C], `dec_w` [C,kD] as 2-D views."""
    _require_cuda_f32(enc_w, embed, dec_w, dec_b, names=("enc.weight", "embed", "dec.weight", "dec.bias"))
    D, M = embed.shape
    C = dec_w.shape[0]
    lib = _capi.load()
    prep = torch.empty((int(lib.ammc_mem_prep_bytes(C, D, M, k)),), dtype=torch.uint8, device=embed.device)
    with torch.cuda.device(embed.device):
        _capi.call("ammc_mem_prepare", _p(enc_w.contiguous()), _p(embed.contiguous()), _p(dec_w.contiguous()),
                   _p(dec_b.contiguous()), _p(prep), prep.numel(), b, h, w, C, D, M, k, _stream())
    _count(8)
    return prep


def mem_forward_raw(x, enc_w, enc_b, embed, dec_w, dec_b, k: int, residual: bool, want_stats: bool,
                    want_planes=False, prep: Optional[torch.Tensor] = None):
    """One fused-module forward.  Returns dict(out, q1[N,D], idx[N,k], z[N,D], sse_frame[b], diff[1], counts, embed_sum).
    want_planes: False | 'bf16' (True) | 'q' -- also emit `out` as the AMFT block's NHWC operand in that format."""
    want_planes = "bf16" if want_planes is True else want_planes
    _require_cuda_f32(x, enc_w, enc_b, embed, dec_w, dec_b, names=("x", "enc.weight", "enc.bias", "embed", "dec.weight", "dec.bias"))
    if x.dim() != 4:
        raise RuntimeError("ammc_b200: memory module input must be [b, C, h, w], got %s" % (tuple(x.shape),))
    _check_device(x.device)
    x = x.contiguous()
    b, C, h, w = x.shape
    D, M = embed.shape
    if enc_w.shape[0] != D or enc_w.shape[1] != C or dec_w.shape[0] != C or dec_w.shape[1] != k * D:
        raise RuntimeError("ammc_b200: inconsistent memory-module parameter shapes")
    N = b * h * w
    dev = x.device
    out = torch.empty_like(x)
    q1 = torch.empty((N, D), dtype=torch.float32, device=dev)
    idx = torch.empty((N, k), dtype=torch.int64, device=dev)
    z = torch.empty((N, D), dtype=torch.float32, device=dev)
    sse = torch.empty((b,), dtype=torch.float32, device=dev)
    diff = torch.empty((1,), dtype=torch.float32, device=dev)
    counts = torch.empty((M,), dtype=torch.float32, device=dev) if want_stats else None
    esum = torch.empty((D, M), dtype=torch.float32, device=dev) if want_stats else None
    lib = _capi.load()
    planes = None
    if want_planes and lib.ammc_mem_dec_uses_tensor(b, h, w, C, D, M, k):
        if want_planes == "q" and C % 256 == 0:
            planes = QPlanes.empty(b, h, w, C, dev)
        else:
            planes = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev)
    ws = _workspace(lib.ammc_mem_workspace_bytes(b, h, w, C, D, M, k), dev)
    with torch.cuda.device(dev):
        _capi.call("ammc_mem_fwd", _p(x), _p(enc_w.contiguous()), _p(enc_b.contiguous()), _p(embed.contiguous()),
                   _p(dec_w.contiguous()), _p(dec_b.contiguous()), _p(out), _p(q1), _p(idx), _p(z), _p(sse), _p(diff),
                   _p(counts), _p(esum), _p(planes), int(isinstance(planes, QPlanes)), _p(prep), _p(ws), ws.numel(), b, h,
                   w, C, D, M, k, int(bool(residual)), _stream())
    # fused front kernel + re-scan + commit + dec when prepared; otherwise + 8 parameter-prep launches (staged kernels: more)
    _count((4 if prep is not None else 12) + (2 if want_stats else 0))
    if planes is not None:
        attach_planes(out, planes)
    return dict(out=out, q1=q1, idx=idx, z=z, sse_frame=sse, diff=diff, counts=counts, embed_sum=esum, x=x)


def mem_forward_io16(x, enc_w, enc_b, embed, dec_w, dec_b, k: int, residual: bool, want_planes=False,
                     prep: Optional[torch.Tensor] = None):
    """bf16 feature-I/O variant of the module forward (BASELINE configs[2]; eval): x, out bf16 NCHW, everything else as
    mem_forward_raw.  Shapes the fused front kernel serves run natively (`ammc_mem_fwd_io16`: x enters the enc GEMM as its
    own bf16 hi plane, residual + bias summed in fp32, one rounding at the store); other shapes widen / narrow around the
    fp32 kernels.  Either way idx, q1, z, diff and the operand planes equal the fp32 path's on the widened input."""
    from .preprocess import widen_bf16, narrow_bf16
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4:
        raise RuntimeError("ammc_b200: mem_forward_io16 needs a CUDA bfloat16 [b, C, h, w] tensor")
    want_planes = "bf16" if want_planes is True else want_planes
    _require_cuda_f32(enc_w, enc_b, embed, dec_w, dec_b, names=("enc.weight", "enc.bias", "embed", "dec.weight", "dec.bias"))
    _check_device(x.device)
    x = x.contiguous()
    b, C, h, w = x.shape
    D, M = embed.shape
    lib = _capi.load()
    if not lib.ammc_mem_io16_supported(b, h, w, C, D, M, k):
        r = mem_forward_raw(widen_bf16(x), enc_w, enc_b, embed, dec_w, dec_b, k, residual, False, want_planes=want_planes,
                            prep=prep)
        out = narrow_bf16(r["out"])
        planes = planes_of(r["out"], "q") or planes_of(r["out"], "bf16")
        if planes is not None:
            attach_planes(out, planes)
        r["out"], r["x"] = out, x
        return r
    if enc_w.shape[0] != D or enc_w.shape[1] != C or dec_w.shape[0] != C or dec_w.shape[1] != k * D:
        raise RuntimeError("ammc_b200: inconsistent memory-module parameter shapes")
    N = b * h * w
    dev = x.device
    out = torch.empty_like(x)
    q1 = torch.empty((N, D), dtype=torch.float32, device=dev)
    idx = torch.empty((N, k), dtype=torch.int64, device=dev)
    z = torch.empty((N, D), dtype=torch.float32, device=dev)
    sse = torch.empty((b,), dtype=torch.float32, device=dev)
    diff = torch.empty((1,), dtype=torch.float32, device=dev)
    planes = None
    if want_planes:
        planes = QPlanes.empty(b, h, w, C, dev) if want_planes == "q" else torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev)
    ws = _workspace(lib.ammc_mem_workspace_bytes(b, h, w, C, D, M, k), dev)
    with torch.cuda.device(dev):
        _capi.call("ammc_mem_fwd_io16", _p(x), _p(enc_w.contiguous()), _p(enc_b.contiguous()), _p(embed.contiguous()),
                   _p(dec_w.contiguous()), _p(dec_b.contiguous()), _p(out), _p(q1), _p(idx), _p(z), _p(sse), _p(diff),
                   _p(planes), int(isinstance(planes, QPlanes)), _p(prep), _p(ws), ws.numel(), b, h, w, C, D, M, k,
                   int(bool(residual)), _stream())
    _count(4 if prep is not None else 12)
    if planes is not None:
        attach_planes(out, planes)
    return dict(out=out, q1=q1, idx=idx, z=z, sse_frame=sse, diff=diff, counts=None, embed_sum=None, x=x)


def check_pipeline_watchdog():
    """Synchronise and raise if a tcgen05 pipeline wait hit its bound (a kernel bug, never expected in production; the
    waiting thread traps, so the failure also surfaces on the next CUDA call of any kind)."""
    rec = (ctypes.c_int * 4)()
    rc = _capi.load().ammc_pipeline_check(rec)
    if rc < 0:
        _capi.check(rc, "ammc_pipeline_check")
    if rc == 1:
        raise RuntimeError("ammc_b200: tcgen05 pipeline wait timed out: family=%d tag=%d block=%d thread=%d"
                           % (rec[0], rec[1], rec[2], rec[3]))


def set_conv_pair_mode(on):
    """True/1: CTA-pair (cta_group::2) conv kernel with fused hi/lo stages (default); 3: pair kernel streaming K three
    times (bit-identical to the single-CTA kernel); False/0: single-CTA kernel.  For A/B measurements only."""
    _capi.call("ammc_set_conv_pair_mode", int(on))


def set_conv_halo_mode(on: bool):
    """True (default): wide 3x3 layers with Cout 64/128 use the halo kernel (tap operands as shifted descriptors over
    one halo tile); False: always the generic implicit GEMM.  For A/B measurements only."""
    _capi.call("ammc_set_conv_halo_mode", int(bool(on)))


def set_addressing_mode(mode: str = "auto"):
    """'auto' | 'fp32' (generic CUDA-core kernel) | 'tensor' (tcgen05 filter + exact refine); process-wide."""
    _capi.call("ammc_set_addressing_mode", {"auto": 0, "fp32": 1, "tensor": 2}[mode])


def last_addressing_stats(device=None):
    """(rows that needed the exact re-scan, path used: 1 = fp32 kernel, 2 = tensor-core filter) of the last memory /
    Quantize_topk forward issued on the current stream of `device`.  Synchronises."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = _workspaces.get((device.index, torch.cuda.current_stream(device).cuda_stream))
    if ws is None:
        return None
    st = ws[:8].view(torch.int32).cpu()
    return int(st[0]), int(st[1])


class MemoryModuleFn(torch.autograd.Function):
    """enc 1x1 -> top-k addressing -> read -> dec 1x1 (+ residual); reference Code/models/unet.py:325-331,384-387.

    forward(x, enc_w[D,C,1,1], enc_b, embed[D,M], dec_w[C,kD,1,1], dec_b, k, residual, want_stats)
      -> out[b,C,h,w], diff[1], q1[b,h,w,D], idx[N,k], sse_frame[b], counts[M]|empty, embed_sum[D,M]|empty
    backward: SURVEY.md Appendix A (no gradient to the bank; the read carries none to z).
    """

    @staticmethod
    def forward(ctx, x, enc_w, enc_b, embed, dec_w, dec_b, k, residual, want_stats, want_planes=False, prep=None):
        # want_planes is decided by the caller: grad mode is always off inside Function.forward
        r = mem_forward_raw(x, enc_w.reshape(enc_w.shape[0], -1), enc_b, embed, dec_w.reshape(dec_w.shape[0], -1),
                            dec_b, k, residual, want_stats, want_planes=want_planes, prep=prep)
        b, C, h, w = r["x"].shape
        D, M = embed.shape
        # training: the caller updates the bank in place right after this forward (EMA, unet.py:298-309);
        # backward needs the PRE-update items, exactly what reference autograd keeps alive -> snapshot (D*M floats)
        bank = embed.clone() if want_stats else embed
        ctx.save_for_backward(r["x"], enc_w, bank, r["idx"], r["z"])
        ctx.dims = (b, h, w, C, D, M, k, bool(residual))
        ctx.wshapes = (enc_w.shape, dec_w.shape)
        empty = x.new_empty((0,))
        outs = (r["out"], r["diff"], r["q1"].view(b, h, w, D), r["idx"], r["sse_frame"],
                r["counts"] if want_stats else empty, r["embed_sum"] if want_stats else empty)
        ctx.mark_non_differentiable(*outs[3:])
        # outputs the loss does not use (q1 in every training script of the reference) arrive as None in backward instead of
        # as zero tensors autograd would have to fill -- [N, D] floats per call for q1 -- and the kernel would have to read
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_out, g_diff, g_q1, *_unused):
        x, enc_w, embed, idx, z = ctx.saved_tensors
        b, h, w, C, D, M, k, residual = ctx.dims
        dev = x.device
        g_out = (torch.zeros_like(x) if g_out is None else g_out.contiguous())
        g_diff = (torch.zeros((1,), dtype=torch.float32, device=dev) if g_diff is None else g_diff.contiguous())
        g_q1 = None if g_q1 is None else g_q1.contiguous()
        _require_cuda_f32(g_out, g_diff, g_q1, names=("grad_out", "grad_diff", "grad_q1"))
        gx = torch.empty_like(x)
        g_enc_w = torch.empty((D, C), dtype=torch.float32, device=dev)
        g_enc_b = torch.empty((D,), dtype=torch.float32, device=dev)
        g_dec_w = torch.empty((C, k * D), dtype=torch.float32, device=dev)
        g_dec_b = torch.empty((C,), dtype=torch.float32, device=dev)
        lib = _capi.load()
        ws = _workspace(lib.ammc_mem_bwd_workspace_bytes(b, h, w, C, D, M, k), dev)
        with torch.cuda.device(dev):
            _capi.call("ammc_mem_bwd", _p(x), _p(enc_w.reshape(D, C).contiguous()), _p(embed.contiguous()), _p(idx),
                       _p(z), _p(g_out), _p(g_diff), _p(g_q1), _p(gx), _p(g_enc_w), _p(g_enc_b), _p(g_dec_w),
                       _p(g_dec_b), _p(ws), ws.numel(), b, h, w, C, D, M, k, int(residual), _stream())
        _count(10)
        es, ds = ctx.wshapes
        return gx, g_enc_w.view(es), g_enc_b, None, g_dec_w.view(ds), g_dec_b, None, None, None, None, None


class QuantizeFn(torch.autograd.Function):
    """Quantize_topk.forward on a [b,h,w,D] input (reference unet.py:282-313).

    -> read[b,h,w,kD], diff (0-d), q1[b,h,w,D], idx[N,k], sse_frame[b], counts, embed_sum
    """

    @staticmethod
    def forward(ctx, z, embed, k, want_stats):
        _require_cuda_f32(z, embed, names=("input", "embed"))
        _check_device(z.device)
        D, M = embed.shape
        if z.shape[-1] != D:
            raise RuntimeError("ammc_b200: Quantize_topk input last dim %d != dim %d" % (z.shape[-1], D))
        zc = z.contiguous()
        lead = zc.shape[:-1]
        N = zc.numel() // D
        frames = lead[0] if len(lead) >= 1 and lead[0] > 0 else 1
        rows = N // frames
        dev = z.device
        read = torch.empty((N, k * D), dtype=torch.float32, device=dev)
        q1 = torch.empty((N, D), dtype=torch.float32, device=dev)
        idx = torch.empty((N, k), dtype=torch.int64, device=dev)
        sse = torch.empty((frames,), dtype=torch.float32, device=dev)
        diff = torch.empty((1,), dtype=torch.float32, device=dev)
        counts = torch.empty((M,), dtype=torch.float32, device=dev) if want_stats else None
        esum = torch.empty((D, M), dtype=torch.float32, device=dev) if want_stats else None
        lib = _capi.load()
        ws = _workspace(lib.ammc_quantize_workspace_bytes(N, D, M, k), dev)
        with torch.cuda.device(dev):
            _capi.call("ammc_quantize_fwd", _p(zc), _p(embed.contiguous()), _p(read), _p(q1), _p(idx), _p(sse),
                       _p(diff), _p(counts), _p(esum), _p(ws), ws.numel(), N, rows, D, M, k, _stream())
        _count(5 + (2 if want_stats else 0))
        ctx.save_for_backward(zc, embed.clone() if want_stats else embed, idx)
        ctx.dims = (N, D, M, k)
        empty = z.new_empty((0,))
        outs = (read.view(*lead, k * D), diff.view(()), q1.view(*lead, D), idx, sse,
                counts if want_stats else empty, esum if want_stats else empty)
        ctx.mark_non_differentiable(outs[0], *outs[3:])
        return outs

    @staticmethod
    def backward(ctx, _g_read, g_diff, g_q1, *_unused):
        zc, embed, idx = ctx.saved_tensors
        N, D, M, k = ctx.dims
        dev = zc.device
        g_diff = (torch.zeros((1,), dtype=torch.float32, device=dev) if g_diff is None
                  else g_diff.reshape(1).contiguous())
        g_q1 = None if g_q1 is None else g_q1.contiguous()
        gz = torch.empty_like(zc)
        lib = _capi.load()
        ws = _workspace(lib.ammc_quantize_bwd_workspace_bytes(N, D, M), dev)
        with torch.cuda.device(dev):
            _capi.call("ammc_quantize_bwd", _p(zc), _p(embed.contiguous()), _p(idx), _p(g_diff), _p(g_q1), _p(gz),
                       _p(ws), ws.numel(), N, D, M, k, _stream())
        _count(3)
        return gz, None, None, None


def embed_code(ids: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """Quantize_topk.embed_code (unet.py:315-316)."""
    _require_cuda_f32(embed, names=("embed",))
    if not ids.is_cuda or ids.dtype != torch.int64:
        raise RuntimeError("ammc_b200: embed_code ids must be a CUDA int64 tensor")
    D, M = embed.shape
    idc = ids.contiguous()
    out = torch.empty((*ids.shape, D), dtype=torch.float32, device=embed.device)
    with torch.cuda.device(embed.device):
        _capi.call("ammc_embed_code", _p(idc), _p(embed.contiguous()), _p(out), idc.numel(), D, M, _stream())
    _count(1)
    return out


def ema_update_(embed, cluster_size, embed_avg, counts, embed_sum, decay: float, eps: float):
    """In-place EMA bank update on the registered buffers (unet.py:298-309)."""
    _require_cuda_f32(embed, cluster_size, embed_avg, counts, embed_sum)
    for t in (embed, cluster_size, embed_avg):
        if not t.is_contiguous():
            raise RuntimeError("ammc_b200: memory-bank buffers must be contiguous")
    D, M = embed.shape
    with torch.cuda.device(embed.device):
        _capi.call("ammc_ema_update", _p(embed), _p(cluster_size), _p(embed_avg), _p(counts.contiguous()),
                   _p(embed_sum.contiguous()), D, M, float(decay), float(eps), _stream())
    _count(2)
    for t in (embed, cluster_size, embed_avg):        # mutated through raw pointers: tell autograd and the caches
        torch.autograd.graph.increment_version(t)     # keyed on tensor versions (prepared parameters, packed weights)


# --------------------------------------------------------------------------------------------------
# AMFT
# --------------------------------------------------------------------------------------------------
def pack_conv_weights(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] (or [Cout, Cin, 1, 1]) fp32 -> [2, Cout, taps*Cin] bf16 (hi, lo planes, k = tap*Cin + cin)."""
    _require_cuda_f32(w, names=("conv weight",))
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    if taps not in (1, 9):
        raise RuntimeError("ammc_b200: only 3x3 and 1x1 conv weights are supported")
    wp = torch.empty((2, Cout, taps * Cin), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_conv_weights" if taps == 9 else "ammc_pack_conv_weights_1x1", _p(w.contiguous()),
                   _p(wp), Cout, Cin, _stream())
    _count(1)
    return wp


def q_conv_supported(Cin: int, Cout: int) -> bool:
    """Shapes the fp16 + e4m3 (precision 2) conv kernel serves; others run precision 3."""
    return Cin % 128 == 0 and Cout % 256 == 0


def pack_conv_weights_q(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] / [Cout, Cin, 1, 1] fp32 -> q weight buffer (uint8; include/ammc_b200.h)."""
    _require_cuda_f32(w, names=("conv weight",))
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    if taps not in (1, 9):
        raise RuntimeError("ammc_b200: only 3x3 and 1x1 conv weights are supported")
    wq = torch.empty((int(_capi.load().ammc_q_weight_bytes(Cout, taps * Cin)),), dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_conv_weights_q", _p(w.contiguous()), _p(wq), Cout, Cin, taps, _stream())
    _count(2)
    return wq


def pack_conv_weights_q_pair(w: torch.Tensor):
    """[Cout, Cin, 3, 3] fp32 -> (q weights of the conv, q weights of its data gradient) with one max|w| reduction."""
    _require_cuda_f32(w, names=("conv weight",))
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    if taps not in (1, 9):
        raise RuntimeError("ammc_b200: only 3x3 and 1x1 conv weights are supported")
    lib = _capi.load()
    wq = torch.empty((int(lib.ammc_q_weight_bytes(Cout, taps * Cin)),), dtype=torch.uint8, device=w.device)
    wd = torch.empty((int(lib.ammc_q_weight_bytes(Cin, taps * Cout)),), dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_conv_weights_q_pair", _p(w.contiguous()), _p(wq), _p(wd), Cout, Cin, taps, _stream())
    _count(3)
    return wq, wd


def pack_nhwc_q(x: torch.Tensor) -> QPlanes:
    """[b, C, h, w] fp32 NCHW -> q operand buffer (max|x| reduction + pack)."""
    _require_cuda_f32(x, names=("activation",))
    b, C, h, w = x.shape
    q = QPlanes.empty(b, h, w, C, x.device)
    with torch.cuda.device(x.device):
        _capi.call("ammc_pack_nhwc_q", _p(x.contiguous()), _p(q.buf), b, C, h, w, _stream())
    _count(3)
    return q


def pack_nhwc_q_planes(x: torch.Tensor):
    """[b, C, h, w] fp32 NCHW -> (QPlanes, bf16 hi/lo planes [2,b,h,w,C]) in one pass over x (C % 8 == 0)."""
    _require_cuda_f32(x, names=("activation",))
    b, C, h, w = x.shape
    q = QPlanes.empty(b, h, w, C, x.device)
    xp = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _capi.call("ammc_pack_nhwc_q_planes", _p(x.contiguous()), _p(q.buf), _p(xp), b, C, h, w, _stream())
    _count(3)
    return q, xp


def bn_fold(gamma, beta, mean, var, eps: float):
    _require_cuda_f32(gamma, beta, mean, var)
    C = gamma.numel()
    scale = torch.empty((C,), dtype=torch.float32, device=gamma.device)
    shift = torch.empty_like(scale)
    with torch.cuda.device(gamma.device):
        _capi.call("ammc_bn_fold", _p(gamma.contiguous()), _p(beta.contiguous()), _p(mean.contiguous()),
                   _p(var.contiguous()), float(eps), _p(scale), _p(shift), C, _stream())
    _count(1)
    return scale, shift


def pack_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[b, C, h, w] fp32 NCHW -> [2, b, h, w, C] bf16 (hi, lo planes)."""
    _require_cuda_f32(x, names=("activation",))
    b, C, h, w = x.shape
    xp = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _capi.call("ammc_pack_nhwc", _p(x.contiguous()), _p(xp), b, C, h, w, _stream())
    _count(1)
    return xp


def conv3x3_bn_relu(xp, wp, scale, shift, *, to_planes: bool, residual: Optional[torch.Tensor] = None,
                    precision: int = 3, relu: bool = True):
    """One implicit-GEMM conv.  precision 1/3: xp [2,b,h,w,Cin] bf16, wp [2,Cout,9Cin] bf16; precision 2: xp a QPlanes,
    wp a q weight buffer.  to_planes=True -> operand planes of the same format ([2,b,h,w,Cout] bf16 or QPlanes);
    else fp32 NCHW [b,Cout,h,w] (+ residual)."""
    if precision == 2:
        if not isinstance(xp, QPlanes):
            raise RuntimeError("ammc_b200: precision 2 takes q-format operands (pack_nhwc_q / pack_conv_weights_q)")
        b, h, w, Cin = xp.shape
        Cout = scale.numel()
        taps = (wp.numel() - 16 - 4 * Cout) // (4 * Cout * Cin)
    else:
        _, b, h, w, Cin = xp.shape
        Cout = wp.shape[1]
        taps = wp.shape[2] // Cin
    dev = xp.device
    _check_device(dev)
    if not to_planes:
        out_p = None
    elif precision == 2:
        out_p = QPlanes.empty(b, h, w, Cout, dev)
    else:
        out_p = torch.empty((2, b, h, w, Cout), dtype=torch.bfloat16, device=dev)
    if residual is not None and residual.dtype == torch.bfloat16 and not to_planes:
        # bf16 feature-I/O variant: bf16 residual in, bf16 NCHW out (fp32 sum, one rounding at the store)
        if not residual.is_cuda:
            raise RuntimeError("ammc_b200: residual must live on a CUDA device (there is no CPU path)")
        if Cout % 256 == 0:
            out16 = torch.empty((b, Cout, h, w), dtype=torch.bfloat16, device=dev)
            conv_layer(xp, wp, scale, shift, taps=taps, act=int(bool(relu)), out_nchw=out16, residual=residual.contiguous(),
                       precision=precision, io_bf16=True)
            return out16
        from .preprocess import widen_bf16, narrow_bf16
        return narrow_bf16(conv3x3_bn_relu(xp, wp, scale, shift, to_planes=False, residual=widen_bf16(residual),
                                           precision=precision, relu=relu))
    out_n = None if to_planes else torch.empty((b, Cout, h, w), dtype=torch.float32, device=dev)
    if residual is not None:
        _require_cuda_f32(residual, names=("residual",))
        residual = residual.contiguous()
    with torch.cuda.device(dev):
        if PROFILE["on"]:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        _capi.call("ammc_conv3x3_bn_relu" if taps == 9 else "ammc_conv1x1_bn_relu", _p(xp), _p(wp), _p(scale), _p(shift), _p(out_p), _p(out_n), _p(residual),
                   b, Cin, Cout, h, w, int(precision), int(bool(relu)), _stream())
        if PROFILE["on"]:
            ev1.record()
            PROFILE["events"].append((ev0, ev1))
    _count(1)
    return out_p if to_planes else out_n


# --------------------------------------------------------------------------------------------------
# general layer form of the conv engine (U-Net encoder / decoder)
# --------------------------------------------------------------------------------------------------
def conv_layer(in_planes, wp, scale, shift, *, taps: int = 9, act: int = 1, Cin: Optional[int] = None, in_c_off: int = 0,
               out_planes=None, out_c_off: int = 0, out_nchw=None, residual=None, cout_valid: int = 0,
               up2x: bool = False, precision: int = 3, io_bf16: bool = False):
    """One launch of `ammc_conv_layer_run`.  io_bf16: out_nchw / residual are bf16 NCHW tensors (CTA-pair kernel only).  in_planes [2,b,h,w,in_cs] bf16 (channel window [in_c_off, +Cin));
    wp [2,Cout,taps*Cin]; out_planes [2,b,ho,wo,out_cs] (written at out_c_off) and/or out_nchw [b,cout_valid,h,w]."""
    q_in = isinstance(in_planes, QPlanes)
    if q_in != (precision == 2):
        raise RuntimeError("ammc_b200: precision 2 <=> q-format input planes")
    if q_in:
        b, h, w, in_cs = in_planes.shape
        Cout = scale.numel()
        Cin = in_cs
    else:
        _, b, h, w, in_cs = in_planes.shape
        Cout = wp.shape[1]
        Cin = wp.shape[2] // taps if Cin is None else Cin
    dev = in_planes.device
    _check_device(dev)
    L = _capi.ConvLayer()
    L.in_planes, L.in_cs, L.in_c_off = in_planes.data_ptr(), in_cs, in_c_off
    L.wp, L.taps = wp.data_ptr(), taps
    L.scale, L.shift, L.act = scale.data_ptr(), shift.data_ptr(), int(act)
    if isinstance(out_planes, QPlanes):
        if out_planes.shape[:3] != (b, h, w) or up2x:
            raise RuntimeError("ammc_b200: q-format out_planes must be [%d,%d,%d,cs]" % (b, h, w))
        L.out_planes, L.out_cs, L.out_c_off, L.out_fmt = out_planes.data_ptr(), out_planes.shape[3], out_c_off, 1
    elif out_planes is not None:
        ho, wo = (2 * h, 2 * w) if up2x else (h, w)
        if tuple(out_planes.shape[:4]) != (2, b, ho, wo) or out_planes.dtype != torch.bfloat16:
            raise RuntimeError("ammc_b200: out_planes must be bf16 [2,%d,%d,%d,cs], got %s" % (b, ho, wo, tuple(out_planes.shape)))
        L.out_planes, L.out_cs, L.out_c_off = out_planes.data_ptr(), out_planes.shape[4], out_c_off
    want = torch.bfloat16 if io_bf16 else torch.float32
    for nm, t in (("out_nchw", out_nchw), ("residual", residual)):
        if t is not None and (not t.is_cuda or t.dtype != want):
            raise RuntimeError("ammc_b200: %s must be a CUDA %s tensor, got %s on %s" % (nm, want, t.dtype, t.device))
    if out_nchw is not None:
        L.out_nchw = out_nchw.data_ptr()
    if residual is not None:
        residual = residual.contiguous()
        L.res_nchw = residual.data_ptr()
    L.io_bf16 = int(bool(io_bf16))
    L.cout_valid = cout_valid
    L.b, L.h, L.w, L.Cin, L.Cout = b, h, w, Cin, Cout
    L.up2x, L.precision, L.in_fmt = int(bool(up2x)), int(precision), int(q_in)
    with torch.cuda.device(dev):
        if PROFILE["on"]:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        _capi.call("ammc_conv_layer_run", ctypes.byref(L), _stream())
        if PROFILE["on"]:
            ev1.record()
            PROFILE["events"].append((ev0, ev1))
            PROFILE["tags"].append(dict(b=b, h=h, w=w, Cin=Cin, Cout=Cout, taps=taps, up2x=bool(up2x), precision=precision))
    _count(1)


def pack_conv_weights_padded(w: torch.Tensor, cout_pad: int, cin_pad: int) -> torch.Tensor:
    """[Cout,Cin,3,3] or [Cout,Cin,1,1] fp32 -> [2,cout_pad,taps*cin_pad] bf16 planes, zero-padded."""
    _require_cuda_f32(w, names=("conv weight",))
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3]
    wp = torch.empty((2, cout_pad, taps * cin_pad), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_conv_weights_padded", _p(w.contiguous()), _p(wp), Cout, Cin, cout_pad, cin_pad, taps, _stream())
    _count(1)
    return wp


def pack_convT_weights(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2) weight [Cin,Cout,2,2] -> [2, 4*Cout, Cin] bf16 planes (row (dy*2+dx)*Cout + co)."""
    _require_cuda_f32(w, names=("ConvTranspose2d weight",))
    Cin, Cout = w.shape[0], w.shape[1]
    if tuple(w.shape[2:]) != (2, 2):
        raise RuntimeError("ammc_b200: only the 2x2 / stride-2 transposed conv of unet.py:46 is supported")
    wp = torch.empty((2, 4 * Cout, Cin), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_convt_weights", _p(w.contiguous()), _p(wp), Cin, Cout, _stream())
    _count(1)
    return wp


def pack_nhwc_padded(x: torch.Tensor, c_pad: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda_f32(x, names=("activation",))
    b, C, h, w = x.shape
    xp = torch.empty((2, b, h, w, c_pad), dtype=torch.bfloat16, device=x.device) if out is None else out
    with torch.cuda.device(x.device):
        _capi.call("ammc_pack_nhwc_padded", _p(x.contiguous()), _p(xp), b, C, c_pad, h, w, _stream())
    _count(1)
    return xp


def maxpool2_planes(in_planes: torch.Tensor, C: int, in_c_off: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MaxPool2d(2) on the channel window [in_c_off, +C) of NHWC planes -> dense [2,b,h/2,w/2,C] planes."""
    _, b, h, w, cs = in_planes.shape
    o = torch.empty((2, b, h // 2, w // 2, C), dtype=torch.bfloat16, device=in_planes.device) if out is None else out
    with torch.cuda.device(in_planes.device):
        _capi.call("ammc_maxpool2_planes", _p(in_planes), cs, in_c_off, _p(o), b, h, w, C, _stream())
    _count(1)
    return o


def unpack_nhwc(planes: torch.Tensor, C: Optional[int] = None, c_off: int = 0) -> torch.Tensor:
    """NHWC hi/lo planes (channel window) -> fp32 NCHW tensor."""
    _, b, h, w, cs = planes.shape
    C = cs - c_off if C is None else C
    x = torch.empty((b, C, h, w), dtype=torch.float32, device=planes.device)
    with torch.cuda.device(planes.device):
        _capi.call("ammc_unpack_nhwc", _p(planes), cs, c_off, _p(x), b, C, h, w, _stream())
    _count(1)
    return x


# optional data-parallel hook: callable(sums[2C] float64 CUDA tensor) -> world size; all-reduces (SUM) the per-channel
# BatchNorm sums in place so that statistics are those of the GLOBAL batch (installed by ammcnet_aaai2021_b200.dist)
BN_SYNC = {"allreduce": None}


def bn_batch_stats(y, gamma, beta, running_mean, running_var, momentum: float, eps: float, training: bool):
    """(scale, shift, mean, invstd) of a BatchNorm2d over y [b,C,h,w]; updates the running stats in place when training."""
    b, C, h, w = y.shape
    dev = y.device
    scale, shift, mean, invstd = (torch.empty((C,), dtype=torch.float32, device=dev) for _ in range(4))
    sync = BN_SYNC["allreduce"] if training else None
    if sync is None:
        ws = _workspace(2 * C * 8, dev)
        with torch.cuda.device(dev):
            _capi.call("ammc_bn_batch_stats", _p(y), _p(gamma.contiguous()), _p(beta.contiguous()), _p(running_mean),
                       _p(running_var), _p(scale), _p(shift), _p(mean), _p(invstd), _p(ws), ws.numel(), b, C, h, w,
                       float(momentum), float(eps), int(bool(training)), _stream())
    else:
        sums = torch.empty((2 * C,), dtype=torch.float64, device=dev)
        args = (_p(y), _p(gamma.contiguous()), _p(beta.contiguous()), _p(running_mean), _p(running_var), _p(scale), _p(shift),
                _p(mean), _p(invstd), _p(sums), sums.numel() * 8, b, C, h, w, float(momentum), float(eps), 1)
        with torch.cuda.device(dev):
            _capi.call("ammc_bn_batch_stats_staged", *args, 1, 0.0, _stream())
            world = int(sync(sums))
            _capi.call("ammc_bn_batch_stats_staged", *args, 2, float(world) * b * h * w, _stream())
    _count(3 if training else 1)
    if training:                                      # running statistics were updated in place by the kernel
        torch.autograd.graph.increment_version(running_mean)
        torch.autograd.graph.increment_version(running_var)
    return scale, shift, mean, invstd


def bn_apply(y, scale, shift, *, relu=True, nhwc=False, nchw=False, f32=False, res=None):
    """relu(y*scale+shift) -> (NHWC bf16 planes | None, NCHW bf16 planes | None, fp32 NCHW (+res) | None)."""
    b, C, h, w = y.shape
    dev = y.device
    o_nhwc = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev) if nhwc else None
    o_nchw = torch.empty((2, b, C, h, w), dtype=torch.bfloat16, device=dev) if nchw else None
    o_f32 = torch.empty_like(y) if f32 else None
    with torch.cuda.device(dev):
        _capi.call("ammc_bn_apply", _p(y), _p(scale), _p(shift), int(bool(relu)), _p(o_nhwc), _p(o_nchw), _p(o_f32),
                   _p(None if res is None else res.contiguous()), b, C, h, w, _stream())
    _count(1)
    return o_nhwc, o_nchw, o_f32


def bn_batch_stats_q(y, gamma, beta, running_mean, running_var, momentum: float, eps: float):
    """Training-mode bn_batch_stats that also leaves, in the returned workspace, the bound behind the q scale of
    act(y*scale+shift) (per-rank statistics).  -> (scale, shift, mean, invstd, workspace)"""
    b, C, h, w = y.shape
    dev = y.device
    scale, shift, mean, invstd = (torch.empty((C,), dtype=torch.float32, device=dev) for _ in range(4))
    ws = torch.empty((int(_capi.load().ammc_bn_q_workspace_bytes(C)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _capi.call("ammc_bn_batch_stats_q", _p(y), _p(gamma.contiguous()), _p(beta.contiguous()), _p(running_mean),
                   _p(running_var), _p(scale), _p(shift), _p(mean), _p(invstd), _p(ws), ws.numel(), b, C, h, w,
                   float(momentum), float(eps), _stream())
    _count(3)
    torch.autograd.graph.increment_version(running_mean)
    torch.autograd.graph.increment_version(running_var)
    return scale, shift, mean, invstd, ws


def bn_apply_q(y, scale, shift, ws, *, relu=True, nhwc=True):
    """relu(y*scale+shift) -> (bf16 hi/lo NHWC planes | None, QPlanes); `ws` from bn_batch_stats_q."""
    b, C, h, w = y.shape
    dev = y.device
    o_nhwc = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev) if nhwc else None
    q = QPlanes.empty(b, h, w, C, dev)
    with torch.cuda.device(dev):
        _capi.call("ammc_bn_apply_q", _p(y), _p(scale), _p(shift), int(bool(relu)), _p(o_nhwc), _p(q), _p(ws), b, C, h, w,
                   _stream())
    _count(1)
    return o_nhwc, q


def bn_backward_q(g, y, scale, shift, mean, invstd, *, relu=True, training=True):
    """Gradient through ReLU+BN -> (g_y as bf16 hi/lo NHWC planes, g_y as QPlanes, g_gamma, g_beta); per-rank statistics."""
    b, C, h, w = y.shape
    dev = y.device
    gy_nhwc = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev)
    gy_q = QPlanes.empty(b, h, w, C, dev)
    gg = torch.empty((C,), dtype=torch.float32, device=dev)
    gb = torch.empty((C,), dtype=torch.float32, device=dev)
    ws = _workspace(int(_capi.load().ammc_bn_q_workspace_bytes(C)), dev)
    with torch.cuda.device(dev):
        _capi.call("ammc_bn_backward_q", _p(g.contiguous()), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd),
                   int(bool(relu)), int(bool(training)), _p(gy_nhwc), _p(gy_q), _p(gg), _p(gb), _p(ws), ws.numel(), b, C, h, w,
                   _stream())
    _count(4)
    return gy_nhwc, gy_q, gg, gb


def bn_backward(g, y, scale, shift, mean, invstd, *, relu=True, training=True):
    """Gradient through ReLU+BN: -> (g_y as NHWC bf16 planes, g_gamma, g_beta)."""
    b, C, h, w = y.shape
    dev = y.device
    gy_nhwc = torch.empty((2, b, h, w, C), dtype=torch.bfloat16, device=dev)
    gy_nchw = None
    gg = torch.empty((C,), dtype=torch.float32, device=dev)
    gb = torch.empty((C,), dtype=torch.float32, device=dev)
    sync = BN_SYNC["allreduce"] if training else None
    if sync is None:
        ws = _workspace(2 * C * 8, dev)
        with torch.cuda.device(dev):
            _capi.call("ammc_bn_backward", _p(g.contiguous()), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd),
                       int(bool(relu)), int(bool(training)), _p(gy_nhwc), _p(gy_nchw), _p(gg), _p(gb), _p(ws), ws.numel(),
                       b, C, h, w, _stream())
    else:
        sums = torch.empty((2 * C,), dtype=torch.float64, device=dev)
        gc = g.contiguous()
        args = (_p(gc), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), int(bool(relu)), 1, _p(gy_nhwc), _p(gy_nchw),
                _p(gg), _p(gb), _p(sums), sums.numel() * 8, b, C, h, w)
        with torch.cuda.device(dev):
            _capi.call("ammc_bn_backward_staged", *args, 1, 0.0, _stream())
            world = int(sync(sums))
            _capi.call("ammc_bn_backward_staged", *args, 2, float(world) * b * h * w, _stream())
    _count(4)
    return gy_nhwc, gg, gb


def pack_planes(x):
    """fp32 tensor -> bf16 hi/lo planes with the same layout: [2, *x.shape]."""
    _require_cuda_f32(x)
    xc = x.contiguous()
    xp = torch.empty((2,) + tuple(xc.shape), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _capi.call("ammc_pack_planes", _p(xc), _p(xp), xc.numel(), _stream())
    _count(1)
    return xp


def pack_conv_weights_dgrad(w):
    """[Cout,Cin,3,3] -> [2, Cin, 9*Cout] planes (taps flipped, channels transposed) for the data-gradient conv."""
    _require_cuda_f32(w)
    Cout, Cin = w.shape[0], w.shape[1]
    wp = torch.empty((2, Cin, 9 * Cout), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _capi.call("ammc_pack_conv_weights_dgrad", _p(w.contiguous()), _p(wp), Cout, Cin, _stream())
    _count(1)
    return wp


def conv3x3_wgrad(gy_nhwc_planes, x_nhwc_planes, precision: int = 3):
    """gw [Cout,Cin,3,3] from the NHWC bf16 planes of the output gradient [2,b,h,w,Cout] and of the conv input."""
    _, b, h, w, Cout = gy_nhwc_planes.shape
    Cin = x_nhwc_planes.shape[4]
    gw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=gy_nhwc_planes.device)
    with torch.cuda.device(gw.device):
        _capi.call("ammc_conv3x3_wgrad", _p(gy_nhwc_planes), _p(x_nhwc_planes), _p(gw), b, Cin, Cout, h, w,
                   int(precision), _stream())
    _count(2)
    return gw


class AmftBranchFn(torch.autograd.Function):
    """out = res + double_conv(u): (conv3x3 -> BN -> ReLU) x 2 with autograd (reference unet.py:8-20, 962-965).

    Training mode uses batch statistics (and updates the running buffers in place); eval mode the running ones.
    forward(u, res, w1, gamma1, beta1, rmean1, rvar1, w2, gamma2, beta2, rmean2, rvar2, training, precision, eps, momentum)
    """

    @staticmethod
    def forward(ctx, u, res, w1, g1, b1, rm1, rv1, w2, g2, b2, rm2, rv2, training, precision, eps, momentum):
        _require_cuda_f32(u, res, w1, w2, names=("input", "residual", "conv.0.weight", "conv.3.weight"))
        u = u.contiguous()
        C = u.shape[1]
        dev = u.device
        one = torch.ones((C,), dtype=torch.float32, device=dev)
        zero = torch.zeros((C,), dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)
        # precision 2 (fp16 + e4m3 operands, two tensor-core pass-equivalents) for the forward and data-gradient convs: in
        # training with per-rank BatchNorm statistics, where the BN passes deliver the planes' scales; the weight gradient
        # keeps its split-bf16 planes.  Otherwise (eval-mode BN under autograd, global-batch BN, small channel counts) x3.
        q = bool(training) and int(precision) == 2 and BN_SYNC["allreduce"] is None and q_conv_supported(C, C)
        prec3 = 3 if int(precision) == 2 else int(precision)
        if q:
            uq, up = pack_nhwc_q_planes(u)
            # the data-gradient weights are packed here too (same max|w| reduction) and kept for backward
            w1q, w1d = pack_conv_weights_q_pair(w1) if need_grad else (pack_conv_weights_q(w1), None)
            w2q, w2d = pack_conv_weights_q_pair(w2) if need_grad else (pack_conv_weights_q(w2), None)
            y1 = conv3x3_bn_relu(uq, w1q, one, zero, to_planes=False, precision=2, relu=False)
            sc1, sh1, mu1, is1, ws1 = bn_batch_stats_q(y1, g1, b1, rm1, rv1, momentum, eps)
            a1_nhwc, a1_q = bn_apply_q(y1, sc1, sh1, ws1, relu=True, nhwc=need_grad)
            y2 = conv3x3_bn_relu(a1_q, w2q, one, zero, to_planes=False, precision=2, relu=False)
        else:
            up = pack_nhwc(u)
            y1 = conv3x3_bn_relu(up, pack_conv_weights(w1), one, zero, to_planes=False, precision=prec3, relu=False)
            sc1, sh1, mu1, is1 = bn_batch_stats(y1, g1, b1, rm1, rv1, momentum, eps, training)
            a1_nhwc, _, _ = bn_apply(y1, sc1, sh1, relu=True, nhwc=True)
            y2 = conv3x3_bn_relu(a1_nhwc, pack_conv_weights(w2), one, zero, to_planes=False, precision=prec3, relu=False)
        sc2, sh2, mu2, is2 = bn_batch_stats(y2, g2, b2, rm2, rv2, momentum, eps, training)
        _, _, out = bn_apply(y2, sc2, sh2, relu=True, f32=True, res=res)
        if need_grad:
            ctx.save_for_backward(up, w1, w2, y1, y2, a1_nhwc, sc1, sh1, mu1, is1, sc2, sh2, mu2, is2, one, zero)
            ctx.cfg = (bool(training), prec3, q)
            ctx.dgrad_q = (w1d, w2d) if q else None          # derived buffers, no autograd history
        return out

    @staticmethod
    def backward(ctx, g_out):
        up, w1, w2, y1, y2, a1_nhwc, sc1, sh1, mu1, is1, sc2, sh2, mu2, is2, one, zero = ctx.saved_tensors
        training, precision, q = ctx.cfg
        g_out = g_out.contiguous()
        if q:
            w1d, w2d = ctx.dgrad_q       # w'[ci][co][tap] = w[co][ci][8 - tap] in the q format, packed in forward
            gy2_nhwc, gy2_q, gg2, gb2 = bn_backward_q(g_out, y2, sc2, sh2, mu2, is2, relu=True, training=training)
            gw2 = conv3x3_wgrad(gy2_nhwc, a1_nhwc, precision)
            g_a1 = conv3x3_bn_relu(gy2_q, w2d, one, zero, to_planes=False, precision=2, relu=False)
            gy1_nhwc, gy1_q, gg1, gb1 = bn_backward_q(g_a1, y1, sc1, sh1, mu1, is1, relu=True, training=training)
            gw1 = conv3x3_wgrad(gy1_nhwc, up, precision)
            g_u = conv3x3_bn_relu(gy1_q, w1d, one, zero, to_planes=False, precision=2, relu=False)
            return (g_u, g_out, gw1, gg1, gb1, None, None, gw2, gg2, gb2, None, None, None, None, None, None)
        gy2_nhwc, gg2, gb2 = bn_backward(g_out, y2, sc2, sh2, mu2, is2, relu=True, training=training)
        gw2 = conv3x3_wgrad(gy2_nhwc, a1_nhwc, precision)
        g_a1 = conv3x3_bn_relu(gy2_nhwc, pack_conv_weights_dgrad(w2), one, zero, to_planes=False, precision=precision,
                               relu=False)
        gy1_nhwc, gg1, gb1 = bn_backward(g_a1, y1, sc1, sh1, mu1, is1, relu=True, training=training)
        gw1 = conv3x3_wgrad(gy1_nhwc, up, precision)
        g_u = conv3x3_bn_relu(gy1_nhwc, pack_conv_weights_dgrad(w1), one, zero, to_planes=False, precision=precision,
                              relu=False)
        return (g_u, g_out, gw1, gg1, gb1, None, None, gw2, gg2, gb2, None, None, None, None, None, None)


# --------------------------------------------------------------------------------------------------
# scoring
# --------------------------------------------------------------------------------------------------
def psnr_per_frame(gen: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """[n, ...] x 2 -> psnr[n]; the batched form of the reference's per-frame psnr_error calls."""
    if gen.is_cuda and gen.dtype == torch.bfloat16 or gt.is_cuda and gt.dtype == torch.bfloat16:
        from .preprocess import widen_bf16           # bf16 frame I/O: widened exactly, PSNR itself in fp32
        gen = widen_bf16(gen) if gen.dtype == torch.bfloat16 else gen
        gt = widen_bf16(gt) if gt.dtype == torch.bfloat16 else gt
    _require_cuda_f32(gen, gt, names=("gen_frames", "gt_frames"))
    if gen.shape != gt.shape:
        raise RuntimeError("ammc_b200: psnr needs equal shapes, got %s vs %s (the reference's broadcasting op-stream "
                           "artefact is not reproduced)" % (tuple(gen.shape), tuple(gt.shape)))
    _check_device(gen.device)
    n = gen.shape[0]
    elems = gen[0].numel() if n > 0 else 1
    g, t = gen.contiguous(), gt.contiguous()
    out = torch.empty((n,), dtype=torch.float32, device=gen.device)
    if n == 0:
        return out
    lib = _capi.load()
    ws = _workspace(lib.ammc_psnr_workspace_bytes(n, elems), gen.device)
    with torch.cuda.device(gen.device):
        _capi.call("ammc_psnr_batch", _p(g), _p(t), _p(out), _p(ws), ws.numel(), n, elems, _stream())
    _count(2)
    return out


def score_reduce_device(img: torch.Tensor, fea: torch.Tensor, offsets: torch.Tensor, lam) -> torch.Tensor:
    """Concatenated per-video records (+ int64 offsets [V+1]) -> regularity scores [T - 4V] (eval_metric.py:405-427)."""
    import numpy as np
    _require_cuda_f32(img, fea, names=("img records", "fea records"))
    V = offsets.numel() - 1
    T = img.numel()
    out = torch.empty((T - 4 * V,), dtype=torch.float32, device=img.device)
    l1, l2 = float(lam[0]), float(lam[1])
    # numpy casts the python doubles (1-lam) and lam to float32 before multiplying a float32 array
    oml1, l1f = float(np.float32(1 - l1)), float(np.float32(l1))
    oml2, l2f = float(np.float32(1 - l2)), float(np.float32(l2))
    lib = _capi.load()
    ws = _workspace(lib.ammc_score_workspace_bytes(T, V), img.device)
    with torch.cuda.device(img.device):
        _capi.call("ammc_score_reduce", _p(img.contiguous()), _p(fea.contiguous()), _p(offsets.contiguous()), V,
                   oml1, l1f, oml2, l2f, _p(out), _p(ws), ws.numel(), T, _stream())
    _count(2)
    return out


def roc_auc_device(scores: torch.Tensor, labels: torch.Tensor, pos_label: int = 0) -> torch.Tensor:
    """ROC-AUC on the GPU (sort + tie-aware rank sum); returns a 0-d float64 device tensor, no host sync."""
    _require_cuda_f32(scores, names=("scores",))
    if not labels.is_cuda or labels.dtype != torch.int8 or labels.numel() != scores.numel():
        raise RuntimeError("ammc_b200: roc_auc labels must be a CUDA int8 tensor of the scores' length")
    T = scores.numel()
    out = torch.empty((1,), dtype=torch.float64, device=scores.device)
    lib = _capi.load()
    ws = _workspace(lib.ammc_auc_workspace_bytes(T), scores.device)
    with torch.cuda.device(scores.device):
        _capi.call("ammc_roc_auc", _p(scores.contiguous()), _p(labels.contiguous()), int(pos_label), _p(out), _p(ws),
                   ws.numel(), T, _stream())
    _count(6)
    return out.reshape(())
