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
