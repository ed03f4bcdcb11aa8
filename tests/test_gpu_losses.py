"""GPU parity: fused image-space generator losses (forward values and gradients) against the oracle / the reference's outputs."""
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
from conftest import load_golden, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_frame_losses_vs_reference_golden():
    c, g = load_golden("losses")
    for name, cs in c["cases"].items():
        gen, gt = synth.frames(cs["seed"], cs["b"], cs["C"], cs["h"], cs["w"])
        gd, td = gen.to(DEV).requires_grad_(True), gt.to(DEV)
        li = A.Intensity_Loss()(gd, td)
        lg = A.Gradient_Loss(channels=cs["C"])(gd, td)
        assert_close(li.detach().cpu(), g[name + "_int"], 1e-5, name + ".int")
        assert_close(lg.detach().cpu(), g[name + "_gd"], 1e-5, name + ".gd")
        gi = torch.autograd.grad(li, gd)[0].cpu()
        gg = torch.autograd.grad(lg, gd)[0].cpu()
        if name + "_g_int" in g:
            assert_close(gi, g[name + "_g_int"], 1e-5, name + ".g_int")
            assert_close(gg, g[name + "_g_gd"], 1e-5, name + ".g_gd")
        else:
            assert_close((gi.double() * gt.double()).sum(), g[name + "_g_int_sum"], 1e-4, name + ".g_int_sum")
            assert_close((gg.double() * gt.double()).sum(), g[name + "_g_gd_sum"], 1e-4, name + ".g_gd_sum")


@pytest.mark.parametrize("b,C,h,w", [(1, 3, 1, 1), (2, 2, 5, 300), (3, 3, 33, 17), (1, 8, 4, 4)])
def test_fused_losses_combined_backward_vs_oracle(b, C, h, w):
    """Both losses weighted like loss_zoo.py:86-87 and differentiated together: one backward kernel."""
    gen, gt = synth.frames(50 + h, b, C, h, w)
    gd, td = gen.to(DEV).requires_grad_(True), gt.to(DEV)
    li, lg = A.frame_losses(gd, td)
    (2.0 * li + 0.5 * lg).backward()
    gr = gen.clone().double().requires_grad_(True)
    (2.0 * O.intensity_loss(gr, gt.double()) + 0.5 * O.gradient_loss(gr, gt.double())).backward()
    assert_close(li.detach().cpu(), O.intensity_loss(gen, gt), 1e-5, "int")
    assert_close(lg.detach().cpu(), O.gradient_loss(gen, gt), 1e-5, "gd")
    assert_close(gd.grad.cpu(), gr.grad, 1e-5, "grad")


def test_losses_refuse_what_they_do_not_cover():
    with pytest.raises(RuntimeError, match="CUDA"):
        A.Intensity_Loss()(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4))
    with pytest.raises(RuntimeError, match="equal shape"):
        A.frame_losses(torch.zeros(1, 3, 4, 4, device=DEV), torch.zeros(1, 3, 4, 5, device=DEV))
    with pytest.raises(RuntimeError, match="1..8 channels"):
        A.frame_losses(torch.zeros(1, 9, 4, 4, device=DEV), torch.zeros(1, 9, 4, 4, device=DEV))
    with pytest.raises(RuntimeError, match="alpha=1"):
        A.Gradient_Loss(alpha=2)


def test_training_objectives_vs_reference_golden():
    """Flow_Loss / Adversarial_Loss / Discriminate_Loss (one fused pass each) and the generator objective Twostream_vq_Loss
    against what the reference itself returned, stored and back-propagated (tests/golden/objectives.npz)."""
    c, g = load_golden("objectives")
    for name, cs in c["cases"].items():
        t = {k: v.to(DEV) for k, v in synth.objective_inputs(cs).items()}
        fp = t["flow_pred"].clone().requires_grad_(True)
        lf = A.Flow_Loss()(fp, t["flow_gt"])
        assert lf.dim() == 0
        assert_close(lf.detach().cpu(), g[name + "_flow"], 1e-5, name + ".flow")
        g_fp = torch.autograd.grad(lf, fp)[0].cpu()
        if name + "_g_flow" in g:
            assert_close(g_fp, g[name + "_g_flow"], 1e-5, name + ".g_flow")
        else:
            assert_close((g_fp.double() * t["flow_gt"].cpu().double()).sum(), g[name + "_g_flow_sum"], 1e-4, name + ".g_flow_sum")
        dg = t["d_gen"].clone().requires_grad_(True)
        la = A.Adversarial_Loss()(dg)
        assert_close(la.detach().cpu(), g[name + "_adv"], 1e-5, name + ".adv")
        assert_close(torch.autograd.grad(la, dg)[0].cpu(), g[name + "_g_adv"], 1e-5, name + ".g_adv")
        dr, df = t["d_real"].clone().requires_grad_(True), t["d_gen"].clone().requires_grad_(True)
        ld = A.Discriminate_Loss()(dr, df)
        assert_close(ld.detach().cpu(), g[name + "_dis"], 1e-5, name + ".dis")
        g_dr, g_df = torch.autograd.grad(ld, (dr, df))
        assert_close(g_dr.cpu(), g[name + "_g_dis_real"], 1e-5, name + ".g_dis_real")
        assert_close(g_df.cpu(), g[name + "_g_dis_fake"], 1e-5, name + ".g_dis_fake")
        # generator objective: same weighted sum, same attributes, same gradients to everything the generator produces
        leaves = {k: t[k].clone().requires_grad_(True) for k in ("rgb_out", "op_out", "latent", "d_gen")}
        fn = A.Twostream_vq_Loss(**c["lambdas"])
        loss = fn(t["flow_pred"], t["flow_gt"], leaves["rgb_out"], t["rgb_tgt"], leaves["op_out"], t["op_tgt"],
                  leaves["latent"], leaves["d_gen"])
        assert_close(loss.detach().cpu().reshape(1), g[name + "_g_loss"], 1e-5, name + ".g_loss")
        for k in ("g_loss", "g_adv_loss", "g_flow_loss", "g_int_loss", "g_gd_loss", "g_int_loss_op", "g_latent_loss"):
            assert isinstance(getattr(fn, k), float)
            assert_close(torch.tensor(getattr(fn, k)), g[name + "_attr_" + k], 1e-5, name + ".attr." + k)
        grads = torch.autograd.grad(loss, list(leaves.values()))
        for (k, leaf), gr in zip(leaves.items(), grads):
            if name + "_dg_" + k in g:
                assert_close(gr.cpu(), g[name + "_dg_" + k], 1e-5, name + ".dg_" + k)
            else:
                assert_close((gr.cpu().double() * t[k].cpu().double()).sum(), g[name + "_dg_" + k + "_sum"], 1e-4,
                             name + ".dg_" + k + "_sum")


@pytest.mark.parametrize("n", [1, 3, 255, 1024, 4097, 3 * 1000 * 1000 + 5])
def test_elementwise_objectives_any_length_and_alignment(n):
    """Lengths around the 16-byte vector width and the grid-stride cap, and views that start off a 16-byte boundary."""
    g = torch.Generator().manual_seed(n)
    a, b = torch.randn(n + 1, generator=g), torch.randn(n + 1, generator=g)
    for off in (0, 1):
        ad, bd = a.to(DEV)[off:off + n].requires_grad_(True), b.to(DEV)[off:off + n].requires_grad_(True)
        ar, br = a[off:off + n].double().requires_grad_(True), b[off:off + n].double().requires_grad_(True)
        for ours, ref in ((A.Flow_Loss()(ad, bd), O.flow_loss(ar, br)), (A.Discriminate_Loss()(ad, bd), O.discriminate_loss(ar, br)),
                          (A.Adversarial_Loss()(ad), O.adversarial_loss(ar))):
            assert_close(ours.detach().cpu(), ref.detach(), 1e-5, "value n=%d off=%d" % (n, off))
            go = torch.autograd.grad(ours, [ad, bd], allow_unused=True)
            gr = torch.autograd.grad(ref, [ar, br], allow_unused=True)
            for x, y in zip(go, gr):
                assert (x is None) == (y is None)
                if x is not None:
                    assert_close(x.cpu(), y, 1e-5, "grad n=%d off=%d" % (n, off))


def test_objectives_refuse_what_they_do_not_cover():
    with pytest.raises(RuntimeError, match="CUDA"):
        A.Flow_Loss()(torch.zeros(4), torch.zeros(4))
    with pytest.raises(RuntimeError, match="equal shape"):
        A.Discriminate_Loss()(torch.zeros(1, 1, 4, 4, device=DEV), torch.zeros(1, 1, 4, 5, device=DEV))
    with pytest.raises(RuntimeError, match="float32"):
        A.Adversarial_Loss()(torch.zeros(4, device=DEV, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="empty"):
        A.Adversarial_Loss()(torch.zeros(0, device=DEV))


@pytest.mark.parametrize("b,h,w,hd,wd", [(1, 8, 8, 3, 3), (2, 33, 17, 5, 4), (3, 64, 48, 34, 34)])
def test_generator_objective_every_component_differentiates(b, h, w, hd, wd):
    """All seven outputs of the fused objective carry gradients (a caller may log or weight them separately), every input that
    asks for a gradient gets one (the flow prediction too), and the generator's (rgb, op) commit-loss tuple is accepted."""
    from ammcnet_aaai2021_b200.losses import GenObjectiveFn
    t = synth.objective_inputs(dict(seed=70 + h, b=b, h=h, w=w, hd=hd, wd=wd))
    lam = dict(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.25, lam_lp_op=2.0)
    keys = ("flow_pred", "rgb_out", "op_out", "d_gen")
    lat2 = torch.tensor([[0.031], [0.012]])                                     # two [1]-shaped commit losses, stacked
    wts = torch.tensor([1.0, 0.3, -0.2, 0.5, 0.7, -0.4, 1.5, 9.0])
    dv = {k: v.to(DEV) for k, v in t.items()}
    for k in keys:
        dv[k].requires_grad_(True)
    lat_d = lat2.to(DEV).requires_grad_(True)
    lams = (lam["lam_adv"], lam["lam_gdl"], lam["lam_flow"], lam["lam_lp"], lam["lam_latent"], lam["lam_lp_op"])
    out = GenObjectiveFn.apply(lams, dv["flow_pred"], dv["flow_gt"], dv["rgb_out"], dv["rgb_tgt"], dv["op_out"], dv["op_tgt"],
                               lat_d.reshape(-1), dv["d_gen"])
    (out * wts.to(DEV)).sum().backward()
    rv = {k: v.double() for k, v in t.items()}
    for k in keys:
        rv[k].requires_grad_(True)
    lat_r = lat2.double().requires_grad_(True)
    loss, parts = O.twostream_vq_loss(lam, rv["flow_pred"], rv["flow_gt"], rv["rgb_out"], rv["rgb_tgt"], rv["op_out"], rv["op_tgt"],
                                      lat_r.sum(), rv["d_gen"])
    order = ("g_loss", "g_adv_loss", "g_flow_loss", "g_int_loss", "g_gd_loss", "g_int_loss_op", "g_latent_loss")
    ref = torch.stack([parts[k].reshape(()) for k in order])
    assert_close(out.detach().cpu()[:7], ref.detach(), 1e-5, "out8")
    assert float(out.detach()[7]) == 0.0
    (ref * wts[:7].double()).sum().backward()
    for k in keys:
        assert_close(dv[k].grad.cpu(), rv[k].grad, 1e-5, "grad " + k)
    assert_close(lat_d.grad.cpu(), lat_r.grad, 1e-5, "grad latent")
    # the module form with the tuple the generator returns as its third output
    fn = A.Twostream_vq_Loss(**lam)
    got = fn(dv["flow_pred"], dv["flow_gt"], dv["rgb_out"], dv["rgb_tgt"], dv["op_out"], dv["op_tgt"],
             (lat_d[0], lat_d[1]), dv["d_gen"])
    assert got.dim() == 0
    assert_close(got.detach().cpu(), loss.detach(), 1e-5, "module g_loss")
    assert abs(fn.g_latent_loss - float(lat2.sum())) < 1e-6
    with pytest.raises(RuntimeError, match="one value"):
        fn(dv["flow_pred"], dv["flow_gt"], dv["rgb_out"], dv["rgb_tgt"], dv["op_out"], dv["op_tgt"], lat_d, dv["d_gen"])


def test_composed_objectives_on_the_kernels():
    """rgb_Loss, rgb_vq_Loss, op_loss, op_vq_Loss, op_loss_v1, op_vq_Loss_v1, Twostream_Loss (loss_zoo.py:64-305) on the fused
    kernels: totals, stored attributes and gradients against the oracle in float64 (table pinned to the reference classes
    in tests/test_host_logic.py)."""
    import objective_checks
    objective_checks.check_composed(DEV, 1e-5)
