"""Slot this package's modules behind the reference's own class names (the drop-in boundary).

    import Code.models.unet as unet
    from ammcnet_aaai2021_b200 import patch
    patch.patch_reference(unet)          # before get_model(const): twostream/UNetMem_v7 now build our modules
    # or, on an already built generator (state_dict keys are identical, so weights carry over):
    patch.swap_modules(model.generator)
"""
from __future__ import annotations

import torch.nn as nn

from . import modules as M

_NAMES = ("Quantize_topk", "enc_quan_dec_topk", "enc_quan_dec_res_topk", "bridge")


def patch_reference(unet_module, utils_module=None):
    """Rebind the hot-path classes inside the reference's `Code.models.unet` (and `psnr_error` in Code.utils.utils)."""
    saved = {}
    for n in _NAMES:
        saved[n] = getattr(unet_module, n)
        setattr(unet_module, n, getattr(M, n))
    if utils_module is not None:
        saved["psnr_error"] = utils_module.psnr_error
        utils_module.psnr_error = M.psnr_error
    return saved


def unpatch_reference(unet_module, saved, utils_module=None):
    for n in _NAMES:
        setattr(unet_module, n, saved[n])
    if utils_module is not None and "psnr_error" in saved:
        utils_module.psnr_error = saved["psnr_error"]


_LOSS_NAMES = ("Flow_Loss", "Intensity_Loss", "Gradient_Loss", "Adversarial_Loss", "Discriminate_Loss", "Twostream_vq_Loss",
               "Twostream_Loss", "rgb_Loss", "rgb_vq_Loss", "op_loss", "op_vq_Loss", "op_loss_v1", "op_vq_Loss_v1")


def patch_reference_losses(loss_zoo_module, losses_utils_module=None):
    """Rebind the training objectives inside the reference's `Code.models.losses.loss_zoo` (which imports the element
    classes by name from `losses_utils`, loss_zoo.py:3-7, and defines `Twostream_vq_Loss`, loss_zoo.py:307) and, optionally,
    inside `losses_utils` itself (`Discriminate_Loss` is constructed from there by the training scripts).  Returns what
    `unpatch_reference_losses` needs."""
    from . import losses as L
    saved = {}
    for mod in (loss_zoo_module, losses_utils_module):
        if mod is None:
            continue
        for n in _LOSS_NAMES:
            if hasattr(mod, n):
                saved[(mod.__name__, n)] = (mod, getattr(mod, n))
                setattr(mod, n, getattr(L, n))
    return saved


def unpatch_reference_losses(saved):
    for (_, n), (mod, cls) in saved.items():
        setattr(mod, n, cls)


def _convert(child: nn.Module):
    name = type(child).__name__
    if name == "enc_quan_dec_res_topk" and not isinstance(child, M.enc_quan_dec_res_topk):
        q = child.quan
        new = M.enc_quan_dec_res_topk(q.enc.in_channels, q.quantize.dim, q.quantize.n_embed, k=q.quantize.k)
    elif name == "enc_quan_dec_topk" and not isinstance(child, M.enc_quan_dec_topk):
        new = M.enc_quan_dec_topk(child.enc.in_channels, child.quantize.dim, child.quantize.n_embed, k=child.quantize.k)
    elif name == "Quantize_topk" and not isinstance(child, M.Quantize_topk):
        new = M.Quantize_topk(child.dim, child.n_embed, decay=child.decay, eps=child.eps, k=child.k)
    elif name == "bridge" and not isinstance(child, M.bridge):
        new = M.bridge(in_c=child.O2F.conv[0].in_channels)
    else:
        return None
    ref_param = next(iter(child.parameters()), None)
    if ref_param is None:
        ref_param = next(iter(child.buffers()))
    new = new.to(ref_param.device)
    new.load_state_dict(child.state_dict(), strict=True)
    new.train(child.training)
    return new


def swap_modules(model: nn.Module) -> nn.Module:
    """Replace every reference hot-path sub-module of `model` in place; returns `model`."""
    for name, child in list(model.named_children()):
        new = _convert(child)
        if new is not None:
            setattr(model, name, new)
        else:
            swap_modules(child)
    return model
