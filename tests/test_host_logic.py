"""CPU: host-side mirror of the reference interface -- state_dict layout, record assembly, patching."""
import os

import numpy as np
import pytest
import torch

import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import scoring
from conftest import load_golden

HAVE_REF = os.path.isdir("/root/reference/Code/models")


def test_twostream_matches_reference_known_answers():
    """25.049029 M parameters (docstring at reference Code/models/unet.py:1268-1275) and 222 state_dict entries."""
    m = A.get_twostream()
    assert sum(p.numel() for p in m.parameters()) == 25049029
    sd = m.state_dict()
    assert len(sd) == 222
    assert tuple(sd["rgb.vq_down3.quan.enc.weight"].shape) == (64, 512, 1, 1)
    assert tuple(sd["op.vq_down3.quan.quantize.embed"].shape) == (64, 256)
    assert tuple(sd["rgb.vq_down3.quan.dec.weight"].shape) == (512, 128, 1, 1)
    assert tuple(sd["bridge.O2F.conv.3.weight"].shape) == (512, 512, 3, 3)
    assert "bridge.F20.conv.4.num_batches_tracked" in sd


def test_single_stream_known_answer():
    """UNetMem_v7(12 -> 3 channels, k=2): 7.805891 M parameters (reference unet.py:1232-1235)."""
    m = A.UNetMem_v7(12, 3, embed_dim=64, n_embed=512, k=2)
    assert sum(p.numel() for p in m.parameters()) == 7805891


def test_record_assembly_matches_reference_records():
    c, g = load_golden("records")
    for v, T in enumerate(c["lengths"]):
        rgb_img, rgb_fea = g[f"rgb_img_pred_records_{v}"], g[f"rgb_fea_comm_records_{v}"]
        n_clips = T - 4
        groups = np.array([rgb_fea[4 + 16 * i] for i in range((n_clips + 15) // 16)], np.float32)
        img, fea = scoring.assemble_video_records(rgb_img[4:], groups, clip_len=5)
        assert np.array_equal(img, rgb_img) and np.array_equal(fea, rgb_fea)
        op_fea = g[f"op_fea_comm_records_{v}"]
        og = np.array([op_fea[3 + 16 * i] for i in range((n_clips + 15) // 16)], np.float32)
        _, fea2 = scoring.assemble_video_records(np.zeros(n_clips, np.float32), og, clip_len=4, tail_copy=True)
        assert np.array_equal(fea2, op_fea)


def test_group_commit_from_frames():
    sse = torch.arange(1, 38, dtype=torch.float32)
    out = scoring.group_commit_from_frames(sse, elems_per_frame=10, group=16)
    exp = [sse[0:16].sum() / 160, sse[16:32].sum() / 160, sse[32:37].sum() / 50]
    assert torch.allclose(out, torch.stack(exp))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_swap_and_patch_against_live_reference():
    import ref_harness
    ref_unet, ref_utils, _ = ref_harness.import_reference()
    g = ref_unet.get_twostream(in_channel=(12, 6), out_channel=(3, 2), embed_dim=64, n_embed=256, k=2)
    ref_sd = {k: v.clone() for k, v in g.state_dict().items()}
    A.swap_modules(g)
    assert isinstance(g.bridge, A.bridge) and isinstance(g.rgb.vq_down3, A.enc_quan_dec_res_topk)
    new_sd = g.state_dict()
    assert list(new_sd) == list(ref_sd)
    assert all(torch.equal(new_sd[k], ref_sd[k]) for k in ref_sd)
    ours = A.get_twostream()
    ours.load_state_dict(ref_sd, strict=True)          # reference checkpoint loads into the standalone host model
    saved = A.patch_reference(ref_unet, ref_utils)
    try:
        g2 = ref_unet.get_twostream(in_channel=(12, 6), out_channel=(3, 2), embed_dim=64, n_embed=256, k=2)
        assert isinstance(g2.bridge, A.bridge) and isinstance(g2.op.vq_down3.quan.quantize, A.Quantize_topk)
        assert list(g2.state_dict()) == list(ref_sd)
        assert ref_utils.psnr_error is A.psnr_error
    finally:
        A.unpatch_reference(ref_unet, saved, ref_utils)
    assert ref_unet.bridge is saved["bridge"]


def test_pixel_discriminator_layout():
    """The configuration the reference trains with (Code/models/__init__.py:123-124,323): 4 convolutions, 128-256-512-1 maps,
    LeakyReLU(0.1); a 256 x 256 frame gives a 34 x 34 map (129 -> 65 -> 33 under 4/2/2, +1 under the 4/1/2 head)."""
    d = A.PixelDiscriminator(3, [128, 256, 512, 512], use_norm=False)
    sd = d.state_dict()
    assert list(sd) == ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.4.weight", "net.4.bias",
                        "net.6.weight", "net.6.bias"]
    assert [tuple(sd[k].shape) for k in list(sd)[::2]] == [(128, 3, 4, 4), (256, 128, 4, 4), (512, 256, 4, 4), (1, 512, 4, 4)]
    assert tuple(d(torch.zeros(1, 3, 256, 256)).shape) == (1, 1, 34, 34)
    n = A.PixelDiscriminator(3, [8, 16, 32, 32], use_norm=True)
    assert [type(m).__name__ for m in n.net] == ["Conv2d", "LeakyReLU", "Conv2d", "LeakyReLU", "BatchNorm2d", "Conv2d",
                                                 "LeakyReLU", "BatchNorm2d", "Conv2d"]


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_pixel_discriminator_against_live_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_pix2pix", "/root/reference/Code/models/pix2pix_networks.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(3)
    x = torch.randn(2, 3, 40, 56)
    for use_norm in (False, True):
        r = ref.PixelDiscriminator(3, [16, 32, 64, 64], use_norm=use_norm)
        o = A.PixelDiscriminator(3, [16, 32, 64, 64], use_norm=use_norm)
        assert list(o.state_dict()) == list(r.state_dict())
        o.load_state_dict(r.state_dict(), strict=True)
        assert torch.equal(o(x), r(x))


def test_objectives_refuse_cpu_tensors():
    """The training objectives have no CPU path either."""
    with pytest.raises(RuntimeError, match="CUDA"):
        A.Flow_Loss()(torch.zeros(2, 2, 4, 4), torch.zeros(2, 2, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        A.Discriminate_Loss()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))
    f = A.Twostream_vq_Loss(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
    assert (f.lam_adv, f.lam_gdl, f.lam_flow, f.lam_lp, f.lam_latent, f.lam_lp_op, f.lam_adv_op) == (0.05, 1.0, 2.0, 1.0, 0.1, 2.0, None)
    assert f.g_loss is None and f.g_latent_loss is None
    z = torch.zeros(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        f(z[:, :2], z[:, :2], z, z, z[:, :2], z[:, :2], torch.zeros(1), torch.zeros(1, 1, 3, 3))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_patch_reference_losses_against_live_reference():
    """The reference's loss modules rebound to this package's classes: the `Twostream_vq_Loss` a training script then builds
    through loss_zoo is ours, with the same constructor keywords and sub-module attribute names; unpatching restores."""
    import sys, types
    import ref_harness
    ref_harness.import_reference()
    stub = types.ModuleType("Code.main.constant_train")
    stub.const = types.SimpleNamespace(gpu_idx="0")
    sys.modules.setdefault("Code.main.constant_train", stub)
    import Code.models.losses.losses_utils as LU
    import Code.models.losses.loss_zoo as LZ
    ref_cls = LZ.Twostream_vq_Loss
    ref_obj = ref_cls(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
    saved = A.patch_reference_losses(LZ, LU)
    try:
        ours = LZ.Twostream_vq_Loss(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
        assert isinstance(ours, A.Twostream_vq_Loss) and LU.Discriminate_Loss is A.Discriminate_Loss
        assert LZ.Flow_Loss is A.Flow_Loss and LZ.Intensity_Loss is A.Intensity_Loss
        ref_attrs = {k for k in vars(ref_obj) if k.startswith(("lam_", "g_"))} | set(dict(ref_obj.named_children()))
        our_attrs = {k for k in vars(ours) if k.startswith(("lam_", "g_"))} | set(dict(ours.named_children()))
        assert ref_attrs <= our_attrs, ref_attrs - our_attrs
    finally:
        A.unpatch_reference_losses(saved)
    assert LZ.Twostream_vq_Loss is ref_cls and LU.Discriminate_Loss is not A.Discriminate_Loss


def test_flownet2sd_layout():
    """'Parameter count = 45,371,666' (reference Code/models/flownet2/FlowNetSD.py:4)."""
    f = A.FlowNet2SD()
    assert sum(p.numel() for p in f.parameters()) == 45371666
    sd = f.state_dict()
    assert tuple(sd["deconv4.0.weight"].shape) == (1026, 256, 4, 4) and tuple(sd["inter_conv2.0.weight"].shape) == (64, 194, 3, 3)
    assert tuple(sd["upsampled_flow6_to_5.weight"].shape) == (2, 2, 4, 4) and "conv6_1.0.bias" in sd
    with torch.no_grad():
        assert tuple(f.eval()(torch.rand(1, 3, 2, 64, 64) * 255).shape) == (1, 2, 64, 64)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_flownet2sd_against_live_reference():
    import ref_harness
    ref_harness.import_reference()
    import Code.models.flownet2.models as RM
    torch.manual_seed(5)
    x = torch.rand(2, 3, 2, 64, 128) * 255
    for bn in (False, True):
        r = RM.FlowNet2SD(batchNorm=bn)
        o = A.FlowNet2SD(batchNorm=bn)
        assert list(o.state_dict()) == list(r.state_dict())
        o.load_state_dict(r.state_dict(), strict=True)
        with torch.no_grad():
            assert torch.equal(o.eval()(x), r.eval()(x))
            for a, b in zip(o.train()(x), r.train()(x)):
                assert torch.equal(a, b)


def _oracle_kernels(monkeypatch):
    """The kernel entry points of losses.py replaced by their oracle formulas (CPU tensors allowed)."""
    import ammc_oracle as O
    from ammcnet_aaai2021_b200 import losses as L
    monkeypatch.setattr(L, "frame_losses", lambda a, b: (O.intensity_loss(a, b), O.gradient_loss(a, b)))
    fakes = {L.OBJ_L1: lambda a, b: O.flow_loss(a, b), L.OBJ_LSGAN_G: lambda a, b: O.adversarial_loss(a),
             L.OBJ_LSGAN_D: lambda a, b: O.discriminate_loss(a, b)}
    monkeypatch.setattr(L.ElemLossFn, "apply", staticmethod(lambda mode, a, b: fakes[mode](a, b)))


def test_composed_objectives_composition(monkeypatch):
    """rgb_Loss ... Twostream_Loss: argument order, weights, attribute names and gradient flow, with the kernels replaced by
    their oracle formulas (the same routine runs on the real kernels in tests/test_gpu_losses.py)."""
    import objective_checks
    _oracle_kernels(monkeypatch)
    objective_checks.check_composed("cpu", 1e-5)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_composed_objectives_table_against_live_reference():
    """The weights / attribute table the two checks above rest on, against the reference classes themselves."""
    import sys, types
    import objective_checks as C
    import ref_harness
    ref_harness.import_reference()
    stub = types.ModuleType("Code.main.constant_train")
    stub.const = types.SimpleNamespace(gpu_idx="0")
    sys.modules.setdefault("Code.main.constant_train", stub)
    import Code.models.losses.loss_zoo as LZ
    cuda_orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, (total_attr, terms, argkeys) in C.TABLE.items():
            t = C.inputs("cpu")
            ref = getattr(LZ, name)(**C.LAM)
            got = ref(*[t[k] for k in argkeys])
            want, attrs = C.expected(name, t)
            assert abs(float(got.detach()) - float(want.detach())) <= 1e-6 * abs(float(want.detach())), name
            for a, v in attrs.items():
                assert abs(getattr(ref, a) - float(v.detach())) <= 1e-6 * max(abs(float(v.detach())), 1e-6), name + "." + a
            ours = getattr(A, name)(**C.LAM)
            assert {k for k in vars(ref) if k.startswith(("lam_", "g_"))} <= {k for k in vars(ours) if k.startswith(("lam_", "g_"))} | {total_attr} | set(attrs)
    finally:
        torch.Tensor.cuda = cuda_orig
