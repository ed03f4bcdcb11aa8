"""Image-space generator losses with the reference's class names (Code/models/losses/losses_utils.py:17-59):
`Intensity_Loss` (channel-wise L2 norm, averaged) and `Gradient_Loss` (|dx| + |dy| of the channel-summed difference).

SURVEY section 8(f) rank 4, the part that needs no external weights.  Both losses of a (prediction, target) pair come from
ONE fused forward kernel and ONE backward kernel (`ammc_frame_losses_fwd/bwd`) instead of ~14 ATen kernels and their
autograd twins; `frame_losses` returns both at once, the classes are thin drop-ins for `loss_zoo.py:37-43`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _capi
from .functions import _check_device, _count, _p, _require_cuda_f32, _stream, _workspace


class FrameLossesFn(torch.autograd.Function):
    """(gen, gt) [n, C, H, W] -> (intensity, gradient) 0-d tensors; gradient flows to `gen` only (targets are data)."""

    @staticmethod
    def forward(ctx, gen, gt):
        _require_cuda_f32(gen, gt, names=("gen_frames", "gt_frames"))
        if gen.shape != gt.shape or gen.dim() != 4:
            raise RuntimeError("ammc_b200: frame losses need two [n, C, H, W] tensors of equal shape, got %s / %s"
                               % (tuple(gen.shape), tuple(gt.shape)))
        _check_device(gen.device)
        gen, gt = gen.contiguous(), gt.contiguous()
        n, C, H, W = gen.shape
        out = torch.empty((2,), dtype=torch.float32, device=gen.device)
        lib = _capi.load()
        ws = _workspace(lib.ammc_frame_losses_workspace_bytes(n, H, W), gen.device)
        with torch.cuda.device(gen.device):
            _capi.call("ammc_frame_losses_fwd", _p(gen), _p(gt), _p(out), _p(ws), ws.numel(), n, C, H, W, _stream())
        _count(2)
        ctx.save_for_backward(gen, gt)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_int, g_gd):
        gen, gt = ctx.saved_tensors
        n, C, H, W = gen.shape
        grad = torch.empty_like(gen)
        g_int = None if g_int is None else g_int.contiguous().float().reshape(1)
        g_gd = None if g_gd is None else g_gd.contiguous().float().reshape(1)
        if g_int is None and g_gd is None:
            return None, None
        with torch.cuda.device(gen.device):
            _capi.call("ammc_frame_losses_bwd", _p(gen), _p(gt), _p(g_int), _p(g_gd), _p(grad), n, C, H, W, _stream())
        _count(1)
        return grad, None


def frame_losses(gen_frames: torch.Tensor, gt_frames: torch.Tensor):
    """-> (Intensity_Loss()(gen, gt), Gradient_Loss(channels=C)(gen, gt)) from one fused pass."""
    return FrameLossesFn.apply(gen_frames, gt_frames)


class Intensity_Loss(nn.Module):
    """losses_utils.py:17-28 with l_num=2 (the only value the reference constructs, loss_zoo.py:38,43)."""

    def __init__(self, l_num=2):
        super().__init__()
        if l_num != 2:
            raise RuntimeError("ammc_b200.Intensity_Loss: only l_num=2 (channel-wise L2 norm) is implemented")
        self.l_num = l_num

    def forward(self, gen_frames, gt_frames):
        return frame_losses(gen_frames, gt_frames)[0]


class Gradient_Loss(nn.Module):
    """losses_utils.py:30-59 with alpha=1; `channels` must equal the frames' channel count (as in the reference, whose
    filter would not match otherwise)."""

    def __init__(self, alpha=1, channels=3):
        super().__init__()
        if alpha != 1:
            raise RuntimeError("ammc_b200.Gradient_Loss: only alpha=1 is implemented")
        self.alpha, self.channels = alpha, channels

    def forward(self, gen_frames, gt_frames):
        if gen_frames.shape[1] != self.channels:
            raise RuntimeError("ammc_b200.Gradient_Loss: built for %d channels, got %d" % (self.channels, gen_frames.shape[1]))
        return frame_losses(gen_frames, gt_frames)[1]
