"""GPU frame / flow preprocessing in front of the generator.

The reference's loaders do, per frame and on CPU workers (Code/dataset/two_stream_dataset.py:72-99): decode, BGR->RGB,
`cv2.resize` to 256x256, `ToTensor`, `Normalize(.5,.5)`; per flow field: `readFlow`, `cv2.resize`, two scalings.  Here the
host only decodes; the decoded uint8 frames (a quarter of the bytes of the fp32 tensors) are uploaded from pinned memory
and everything after the decode runs in one kernel per batch, bit-exact with the reference's loaders
(tests/golden/preprocess.npz).  Decoding (TurboJPEG / np.fromfile) stays host code.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _capi
from .functions import _check_device, _count, _p, _stream


def preprocess_frames(frames_bgr_u8: torch.Tensor, size: Tuple[int, int] = (256, 256)) -> torch.Tensor:
    """Decoded frames uint8 [n, h0, w0, 3] (BGR, as cv2.imread / TurboJPEG return them) on a CUDA device ->
    float32 [n, 3, H, W] in (-1, 1); `size` = (width, height) like cv2.resize and the reference's `img_size`."""
    t = frames_bgr_u8
    if not t.is_cuda or t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
        raise RuntimeError("ammc_b200: preprocess_frames needs a CUDA uint8 tensor [n, h, w, 3], got %s %s on %s"
                           % (t.dtype, tuple(t.shape), t.device))
    _check_device(t.device)
    W, H = int(size[0]), int(size[1])
    n, h0, w0, _ = t.shape
    out = torch.empty((n, 3, H, W), dtype=torch.float32, device=t.device)
    if n == 0:
        return out
    with torch.cuda.device(t.device):
        _capi.call("ammc_preprocess_frames_u8", _p(t.contiguous()), _p(out), n, h0, w0, H, W, _stream())
    _count(1)
    return out


def preprocess_flow(flow: torch.Tensor, size: Tuple[int, int] = (256, 256)) -> torch.Tensor:
    """.flo payloads float32 [n, h0, w0, 2] on a CUDA device -> float32 [n, 2, H, W] exactly as `_load_op` produces them
    (channel 1 is derived from channel 0, two_stream_dataset.py:94-95)."""
    t = flow
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4 or t.shape[3] != 2:
        raise RuntimeError("ammc_b200: preprocess_flow needs a CUDA float32 tensor [n, h, w, 2], got %s %s on %s"
                           % (t.dtype, tuple(t.shape), t.device))
    _check_device(t.device)
    W, H = int(size[0]), int(size[1])
    n, h0, w0, _ = t.shape
    out = torch.empty((n, 2, H, W), dtype=torch.float32, device=t.device)
    if n == 0:
        return out
    with torch.cuda.device(t.device):
        _capi.call("ammc_preprocess_flow", _p(t.contiguous()), _p(out), n, h0, w0, H, W, _stream())
    _count(1)
    return out


def widen_bf16(t: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """bf16 CUDA tensor -> fp32 tensor of the same shape (exact); the device side of a bf16 host boundary."""
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise RuntimeError("ammc_b200: widen_bf16 needs a CUDA bfloat16 tensor, got %s on %s" % (t.dtype, t.device))
    _check_device(t.device)
    tc = t.contiguous()
    out = torch.empty(tc.shape, dtype=torch.float32, device=t.device) if out is None else out
    if tc.numel():
        with torch.cuda.device(t.device):
            _capi.call("ammc_cast_bf16_f32", _p(tc), _p(out), tc.numel(), _stream())
        _count(1)
    return out


def narrow_bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32 CUDA tensor -> bf16 tensor of the same shape (round to nearest even)."""
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("ammc_b200: narrow_bf16 needs a CUDA float32 tensor, got %s on %s" % (t.dtype, t.device))
    _check_device(t.device)
    tc = t.contiguous()
    out = torch.empty(tc.shape, dtype=torch.bfloat16, device=t.device)
    if tc.numel():
        with torch.cuda.device(t.device):
            _capi.call("ammc_cast_f32_bf16", _p(tc), _p(out), tc.numel(), _stream())
        _count(1)
    return out


def upload(arrays: Sequence, device) -> torch.Tensor:
    """Stack equally shaped host arrays (decoded frames or flow payloads) in pinned memory and copy them to `device`
    asynchronously on the current stream."""
    import numpy as np
    host = torch.from_numpy(np.stack([np.ascontiguousarray(a) for a in arrays])).pin_memory()
    return host.to(device, non_blocking=True)


def load_video(frames_bgr_u8: Sequence, flows: Sequence, device, size: Tuple[int, int] = (256, 256)):
    """Decoded frames + flow payloads of one sub-video -> (rgb_frames [T,3,H,W], op_frames [T-1,2,H,W]) on `device`, the
    inputs of `VideoScorer.score_video`."""
    return (preprocess_frames(upload(frames_bgr_u8, device), size), preprocess_flow(upload(flows, device), size))
