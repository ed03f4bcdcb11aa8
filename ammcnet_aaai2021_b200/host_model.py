"""Host-side two-stream generator wiring (row a8 of SURVEY.md section 8): the caller of the hot path.

The encoder / decoder layers are stock torch.nn modules with the reference's names (reference Code/models/unet.py
`inconv/down/up` 23-59, `UNetMem_v7` 908-937, `twostream` 967-1007), so a reference checkpoint loads with strict=True;
this package's memory modules and AMFT block sit in the starred region (unet.py:985-994).  The file exists so the
package can be exercised end to end without the reference tree (absent on the GPU box).

Two execution routes for `twostream.forward`:
* training / autograd: the torch.nn layers run on cuDNN exactly as in the reference (north-star: "the conv
  encoder/decoder (left on cuDNN) ... unchanged"), the path's modules run this package's kernels;
* eval + no_grad on a CUDA device (`engine = "tcgen05"`, the default): the whole forward runs on the tcgen05 conv engine
  through `generator.GeneratorEngine`.  Set `model.engine = "cudnn"` to keep the layers on
  cuDNN.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modules import bridge, double_conv, enc_quan_dec_res_topk


class inconv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = double_conv(in_ch, out_ch)

    def forward(self, x):
        return self.conv.conv(x)


class down(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), double_conv(in_ch, out_ch))

    def forward(self, x):
        return self.mpconv[1].conv(self.mpconv[0](x))


class up(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_ch, in_ch // 2, 2, stride=2)
        self.conv = double_conv(in_ch, out_ch)

    def forward(self, x1, x2):
        x1 = self.up(x1)
        dy, dx = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
        x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return self.conv.conv(torch.cat([x2, x1], dim=1))


class UNetMem_v7(nn.Module):
    def __init__(self, input_channels=3, output_channel=3, embed_dim=64, n_embed=512, k=1,
                 layer_nums=4, features_root=64):
        super().__init__()
        self.inc = inconv(input_channels, 64)
        self.down1 = down(64, 128)
        self.down2 = down(128, 256)
        self.down3 = down(256, 512)
        self.up1 = up(512, 256)
        self.up2 = up(256, 128)
        self.up3 = up(128, 64)
        self.outc = nn.Conv2d(64, output_channel, kernel_size=3, padding=1)
        self.vq_down3 = enc_quan_dec_res_topk(512, embed_dim, n_embed, k=k)

    def encode(self, x):
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        return x1, x2, x3, self.down3(x3)

    def decode(self, x4, x3, x2, x1):
        return self.outc(self.up3(self.up2(self.up1(x4, x3), x2), x1))

    def forward(self, x):
        x1, x2, x3, x4 = self.encode(x)
        x4, diff, q1 = self.vq_down3(x4)
        return torch.tanh(self.decode(x4, x3, x2, x1)), diff, q1


class twostream(nn.Module):
    def __init__(self, rgb_in_c, rgb_out_c, op_in_c, op_out_c, embed_dim=64, n_embed=512, k=1,
                 layer_nums=4, features_root=64):
        super().__init__()
        self.rgb = UNetMem_v7(rgb_in_c, rgb_out_c, embed_dim, n_embed, k, layer_nums, features_root)
        self.op = UNetMem_v7(op_in_c, op_out_c, embed_dim, n_embed, k, layer_nums, features_root)
        self.bridge = bridge(in_c=512)
        self.engine = "tcgen05"
        object.__setattr__(self, "_engine_obj", None)

    def forward(self, rgb_x, op_x):
        if (self.engine == "tcgen05" and not self.training and not torch.is_grad_enabled() and rgb_x.is_cuda
                and rgb_x.shape[2] % 8 == 0 and rgb_x.shape[3] % 8 == 0):
            if self._engine_obj is None:
                from .generator import GeneratorEngine
                object.__setattr__(self, "_engine_obj", GeneratorEngine(self))
            return self._engine_obj(rgb_x, op_x)
        r1, r2, r3, r4 = self.rgb.encode(rgb_x)
        self.quant_befor = r4                                   # side attributes the reference keeps (unet.py:986,988)
        r4, rgb_diff, rgb_q = self.rgb.vq_down3(r4)
        self.quant_after = r4
        o1, o2, o3, o4 = self.op.encode(op_x)
        o4, op_diff, op_q = self.op.vq_down3(o4)
        r4, o4 = self.bridge(r4, o4)
        rgb_y = self.rgb.decode(r4, r3, r2, r1)
        op_y = self.op.decode(o4, o3, o2, o1)
        return torch.tanh(rgb_y), torch.tanh(op_y), (rgb_diff, op_diff), (rgb_q, op_q)


def get_twostream(in_channel=(12, 6), out_channel=(3, 2), embed_dim=64, n_embed=256, k=2):
    """Shipped configuration (reference Code/models/unet.py:1241-1249, net_params/*.pkl)."""
    return twostream(in_channel[0], out_channel[0], in_channel[1], out_channel[1], embed_dim=embed_dim,
                     n_embed=n_embed, k=k)


class PixelDiscriminator(nn.Module):
    """Host-side mirror of the reference discriminator (Code/models/pix2pix_networks.py:580-631; built as
    `PixelDiscriminator(3, [128, 256, 512, 512], use_norm=False)`, Code/models/__init__.py:123-124,323): a stack of 4x4
    stride-2 convolutions with LeakyReLU(0.1), optional norm layers after the inner activations, and a 4x4 stride-1 head
    giving one map value per receptive field.  Same constructor arguments and the same `net.<i>` parameter names, so a
    reference discriminator checkpoint loads with strict=True.  The layers are stock torch.nn modules (cuDNN): this is
    training-step plumbing outside the hot path -- what this package adds to the adversarial
    part of the step are the fused objectives in `losses.py` that consume its maps."""

    def __init__(self, input_nc, num_filters, use_norm=False, norm_layer=nn.BatchNorm2d):
        super().__init__()
        inner = getattr(norm_layer, "func", norm_layer)
        bias = (inner != nn.InstanceNorm2d) if use_norm else True
        layers = [nn.Conv2d(input_nc, num_filters[0], kernel_size=4, stride=2, padding=2), nn.LeakyReLU(0.1, True)]
        for cin, cout in zip(num_filters[:-2], num_filters[1:-1]):
            layers += [nn.Conv2d(cin, cout, 4, 2, 2, bias=bias), nn.LeakyReLU(0.1, True)]
            if use_norm:
                layers.append(norm_layer(cout))
        layers.append(nn.Conv2d(num_filters[-1], 1, 4, 1, 2))
        self.net = nn.Sequential(*layers)

    def forward(self, input):
        return self.net(input)


def _flow_conv(bn, cin, cout, stride=1, act=True):
    layers = [nn.Conv2d(cin, cout, 3, stride, 1, bias=not (bn and act))]
    if bn:
        layers.append(nn.BatchNorm2d(cout))
    if act:
        layers.append(nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(*layers)


class FlowNet2SD(nn.Module):
    """Host-side mirror of the frozen flow estimator of the training step (reference Code/models/flownet2/models.py:9-58 on
    Code/models/flownet2/FlowNetSD.py:7-100 and submodules.py:8-48; called twice per step on (last input frame, prediction)
    and (last input frame, target), train_helper.py:309-322).  Same constructor, same module names -- a reference /
    NVIDIA FlowNet2-SD checkpoint loads with strict=True (45 371 666 parameters) -- and the same forward: per-sample channel
    means removed, /255, the two frames stacked on the channel axis, 13 encoder convolutions (6 of them stride 2), a
    4-level refinement decoder (transposed convolutions + flow up-sampling + skip concatenation), x`div_flow` and 4x bilinear
    up-sampling in eval mode; the five multi-scale flows in training mode.  Stock torch.nn layers (cuDNN): plumbing for
    `tools/train_step.py --gan --flownet`, outside the hot path; the `Flow_Loss` that consumes its two outputs is a kernel of
    this package (`losses.py`).  No weights ship with either repository: tools build it with random initial weights."""

    _ENCODER = (("conv0", 6, 64, 1), ("conv1", 64, 64, 2), ("conv1_1", 64, 128, 1), ("conv2", 128, 128, 2),
                ("conv2_1", 128, 128, 1), ("conv3", 128, 256, 2), ("conv3_1", 256, 256, 1), ("conv4", 256, 512, 2),
                ("conv4_1", 512, 512, 1), ("conv5", 512, 512, 2), ("conv5_1", 512, 512, 1), ("conv6", 512, 1024, 2),
                ("conv6_1", 1024, 1024, 1))
    _SKIP = {2: 128, 3: 256, 4: 512, 5: 512}                      # channels of the encoder output joined at each level
    _DECODER = {5: 512, 4: 256, 3: 128, 2: 64}                    # channels the level's transposed convolution produces

    def __init__(self, batchNorm=False, div_flow=20):
        super().__init__()
        self.batchNorm, self.rgb_max, self.div_flow = batchNorm, 255., div_flow
        for name, cin, cout, stride in self._ENCODER:
            setattr(self, name, _flow_conv(batchNorm, cin, cout, stride))
        levels = (5, 4, 3, 2)
        joined = {lvl: self._SKIP[lvl] + self._DECODER[lvl] + 2 for lvl in levels}      # skip + decoded + up-sampled flow
        for lvl in levels:                                          # registration order = the reference's state_dict order
            below = 1024 if lvl == 5 else joined[lvl + 1]
            setattr(self, "deconv%d" % lvl, nn.Sequential(nn.ConvTranspose2d(below, self._DECODER[lvl], 4, 2, 1),
                                                          nn.LeakyReLU(0.1, inplace=True)))
        for lvl in levels:
            setattr(self, "inter_conv%d" % lvl, _flow_conv(batchNorm, joined[lvl], self._DECODER[lvl], act=False))
        self.predict_flow6 = nn.Conv2d(1024, 2, 3, 1, 1)
        for lvl in levels:
            setattr(self, "predict_flow%d" % lvl, nn.Conv2d(self._DECODER[lvl], 2, 3, 1, 1))
        for lvl in levels:
            setattr(self, "upsampled_flow%d_to_%d" % (lvl + 1, lvl), nn.ConvTranspose2d(2, 2, 4, 2, 1))
        for m in self.modules():                                    # FlowNetSD.py:46-56
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                if m.bias is not None:
                    nn.init.uniform_(m.bias)
                nn.init.xavier_uniform_(m.weight)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def forward(self, inputs):
        """inputs [b, 3, 2, h, w] in (0, 255), h and w multiples of 64."""
        mean = inputs.flatten(2).mean(dim=-1)[:, :, None, None, None]
        x = ((inputs - mean) / self.rgb_max).transpose(1, 2).flatten(1, 2)          # frame 0's channels, then frame 1's
        skips = {}
        for name, _, _, _ in self._ENCODER:
            x = getattr(self, name)(x)
            if name.endswith("_1"):
                skips[int(name[4])] = x
        joined = skips[6]
        flows = [self.predict_flow6(joined)]
        for lvl in (5, 4, 3, 2):
            up = getattr(self, "upsampled_flow%d_to_%d" % (lvl + 1, lvl))(flows[-1])
            joined = torch.cat((skips[lvl], getattr(self, "deconv%d" % lvl)(joined), up), 1)
            flows.append(getattr(self, "predict_flow%d" % lvl)(getattr(self, "inter_conv%d" % lvl)(joined)))
        if self.training:
            return tuple(reversed(flows))                           # flow2 ... flow6
        return self.upsample1(flows[-1] * self.div_flow)
