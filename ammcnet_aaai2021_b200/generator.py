"""Eval-mode forward of the whole `twostream` generator on the tcgen05 conv engine.

The reference runs the U-Net encoder/decoder around the memory path as ~40 cuDNN / ATen calls per stream
(Code/models/unet.py:981-1007).  Here every layer is one launch of the same implicit-GEMM kernel that serves the
AMFT block (`ammc_conv_layer_run`), and no tensor between two layers is ever fp32 or NCHW:

* activations stay NHWC bf16 hi/lo planes (the operand format of the engine, fp32-parity with precision=3);
* BatchNorm is folded to scale/shift and applied with ReLU in the producing conv's epilogue (unet.py:11-16);
* `torch.cat([skip, up], 1)` (unet.py:58) never copies: the skip's producer writes channels [0, C) and the transposed
  conv's epilogue scatters into channels [C, 2C) of one pre-allocated buffer;
* `ConvTranspose2d(2, stride=2)` (unet.py:46) is a 1x1 GEMM with 4*Cout columns and a scattering epilogue;
* `outc` + `tanh` (unet.py:918,936-937) is a zero-padded 64-column conv whose epilogue applies bias + tanh and writes
  the 3 (2) real channels as fp32 NCHW;
* the 12- / 6-channel network inputs are zero-padded to 64 channels once when they are packed.

Only inference is covered (frozen BN, no autograd); `host_model.twostream` keeps the cuDNN layers for training.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import functions as F_


def _planes(b, h, w, c, dev):
    return torch.empty((2, b, h, w, c), dtype=torch.bfloat16, device=dev)


class _StreamPack:
    """Packed weights of one UNetMem_v7 stream."""

    def __init__(self, u):
        self.layers: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        dev = u.outc.weight.device
        for name, dc in (("inc", u.inc.conv), ("down1", u.down1.mpconv[1]), ("down2", u.down2.mpconv[1]),
                         ("down3", u.down3.mpconv[1]), ("up1", u.up1.conv), ("up2", u.up2.conv), ("up3", u.up3.conv)):
            for i, (ci, bi) in enumerate(((0, 1), (3, 4))):
                conv, bn = dc.conv[ci], dc.conv[bi]
                cout, cin = conv.weight.shape[0], conv.weight.shape[1]
                cin_pad = (cin + 63) // 64 * 64
                wp = (F_.pack_conv_weights(conv.weight.detach()) if cin_pad == cin
                      else F_.pack_conv_weights_padded(conv.weight.detach(), cout, cin_pad))
                scale, shift = F_.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
                self.layers[f"{name}.{i}"] = (wp, scale, shift)
        for name, upm in (("up1", u.up1.up), ("up2", u.up2.up), ("up3", u.up3.up)):
            cout = upm.weight.shape[1]
            if tuple(upm.kernel_size) != (2, 2) or tuple(upm.stride) != (2, 2):
                raise RuntimeError("ammc_b200 generator engine: `up` must be ConvTranspose2d(2, stride=2)")
            wp = F_.pack_convT_weights(upm.weight.detach())
            bias = upm.bias.detach() if upm.bias is not None else torch.zeros(cout, device=dev)
            self.layers[name + ".up"] = (wp, torch.ones(4 * cout, device=dev), bias.float().repeat(4).contiguous())
        oc = u.outc
        self.out_c = oc.weight.shape[0]
        wp = F_.pack_conv_weights_padded(oc.weight.detach(), 64, 64)
        shift = torch.zeros(64, device=dev)
        if oc.bias is not None:
            shift[: self.out_c] = oc.bias.detach()
        self.layers["outc"] = (wp, torch.ones(64, device=dev), shift)


class GeneratorEngine:
    """`GeneratorEngine(model)(rgb_x, op_x)` == `model(rgb_x, op_x)` of an eval-mode `twostream` (unet.py:981-1007).

    `model` is any module with the reference's attribute layout (rgb / op : UNetMem_v7 with this package's
    `vq_down3`, bridge : this package's `bridge`).  precision 3 = split-bf16 x3 (fp32 parity), 1 = single bf16 pass.
    """

    def __init__(self, model, precision: int = 3):
        self.model = model
        self.precision = precision
        self._key = None
        self._packs = None

    def __getattr__(self, name):
        # attribute passthrough (rgb, op, bridge, ...) so the engine can stand in for the module, e.g. in VideoScorer
        model = self.__dict__.get("model")
        if model is None or name.startswith("__"):
            raise AttributeError(name)
        return getattr(model, name)

    # -- weights ----------------------------------------------------------------------------------
    def _ensure_packed(self):
        tensors = list(self.model.parameters()) + list(self.model.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self._key:
            with torch.no_grad():
                self._packs = {"rgb": _StreamPack(self.model.rgb), "op": _StreamPack(self.model.op)}
            self._key = key

    # -- pieces -----------------------------------------------------------------------------------
    def _conv(self, pk, name, src, **kw):
        wp, scale, shift = pk.layers[name]
        F_.conv_layer(src, wp, scale, shift, precision=self.precision, **kw)

    def _encode(self, pk, x):
        b, _, H, W = x.shape
        dev = x.device
        xin = F_.pack_nhwc_padded(x, 64)
        cats = []
        src, c = xin, 64
        for lvl, name in enumerate(("inc", "down1", "down2")):
            h, w = H >> lvl, W >> lvl
            mid = _planes(b, h, w, c, dev)
            self._conv(pk, name + ".0", src, out_planes=mid)
            cat = _planes(b, h, w, 2 * c, dev)                 # [skip | upsampled] of the matching `up` block
            self._conv(pk, name + ".1", mid, out_planes=cat)   # skip -> channels [0, c)
            cats.append(cat)
            src = F_.maxpool2_planes(cat, c)
            c *= 2
        h, w = H >> 3, W >> 3
        mid = _planes(b, h, w, c, dev)
        self._conv(pk, "down3.0", src, out_planes=mid)
        x4 = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        self._conv(pk, "down3.1", mid, out_nchw=x4)            # the memory module's `enc` reads fp32 NCHW
        return cats, x4

    def _amft_branch(self, dc, src_planes, residual, prec):
        q = prec == 2
        w1, s1, b1 = dc.packed(0, 1, q)
        w2, s2, b2 = dc.packed(3, 4, q)
        b, h, w, c = src_planes.shape if q else src_planes.shape[1:]
        mid = F_.QPlanes.empty(b, h, w, c, src_planes.device) if q else _planes(b, h, w, c, src_planes.device)
        F_.conv_layer(src_planes, w1, s1, b1, out_planes=mid, precision=prec)
        out = _planes(b, h, w, c, src_planes.device)           # up1's transposed conv reads bf16 hi/lo planes
        F_.conv_layer(mid, w2, s2, b2, out_planes=out, residual=residual, precision=prec)
        return out

    def _decode(self, pk, x4p, cats):
        b = x4p.shape[1]
        dev = x4p.device
        src = x4p
        for name, cat in (("up1", cats[2]), ("up2", cats[1]), ("up3", cats[0])):
            c = cat.shape[4] // 2
            self._conv(pk, name + ".up", src, taps=1, act=0, up2x=True, out_planes=cat, out_c_off=c)
            h, w = cat.shape[2], cat.shape[3]
            mid = _planes(b, h, w, c, dev)
            self._conv(pk, name + ".0", cat, out_planes=mid)
            src = _planes(b, h, w, c, dev)
            self._conv(pk, name + ".1", mid, out_planes=src)
        y = torch.empty((b, pk.out_c, src.shape[2], src.shape[3]), dtype=torch.float32, device=dev)
        self._conv(pk, "outc", src, act=2, out_nchw=y, cout_valid=pk.out_c)
        return y

    # -- forward ----------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, rgb_x, op_x):
        m = self.model
        if m.training:
            raise RuntimeError("ammc_b200 generator engine: inference only (call model.eval()); training uses the "
                               "autograd path of host_model.twostream")
        F_._require_cuda_f32(rgb_x, op_x, names=("rgb_x", "op_x"))
        if rgb_x.shape[2] % 8 or rgb_x.shape[3] % 8 or rgb_x.shape[2:] != op_x.shape[2:] or rgb_x.shape[0] != op_x.shape[0]:
            raise RuntimeError("ammc_b200 generator engine: frame height/width must be multiples of 8 and equal for "
                               "both streams, got %s / %s" % (tuple(rgb_x.shape), tuple(op_x.shape)))
        self._ensure_packed()
        pr, po = self._packs["rgb"], self._packs["op"]
        def stream_front(pk, unet, x):                         # encoder + memory module of one stream
            cats, x4 = self._encode(pk, x)
            return (cats,) + tuple(unet.vq_down3(x4)) + (x4,)

        # the appearance and motion streams are independent up to the AMFT block and again after it (unet.py:981-1003)
        (cats_r, r4, rgb_diff, rgb_q, r4_in), (cats_o, o4, op_diff, op_q, _) = F_.concurrently(
            lambda: stream_front(pr, m.rgb, rgb_x), lambda: stream_front(po, m.op, op_x))
        m.quant_befor, m.quant_after = r4_in, r4             # side attributes of the reference forward (unet.py:986,988)
        prec = m.bridge.eval_precision(r4.shape[1])
        fmt, pack = ("q", F_.pack_nhwc_q) if prec == 2 else ("bf16", F_.pack_nhwc)
        px, py = F_.planes_of(r4, fmt), F_.planes_of(o4, fmt)
        px = pack(r4) if px is None else px
        py = pack(o4) if py is None else py
        r4p, o4p = F_.concurrently(lambda: self._amft_branch(m.bridge.O2F, py, r4, prec),     # x' = zx + O2F(zy)   (unet.py:963)
                                   lambda: self._amft_branch(m.bridge.F20, px, o4, prec))     # y' = zy + F20(zx)   (unet.py:964)
        rgb_y, op_y = F_.concurrently(lambda: self._decode(pr, r4p, cats_r), lambda: self._decode(po, o4p, cats_o))
        return rgb_y, op_y, (rgb_diff, op_diff), (rgb_q, op_q)
