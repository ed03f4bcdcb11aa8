"""Shared body of the checks on the composed training objectives (rgb_Loss, rgb_vq_Loss, op_loss, op_vq_Loss, op_loss_v1,
op_vq_Loss_v1, Twostream_Loss -- reference Code/models/losses/loss_zoo.py:64-305).  The same routine runs on the GPU (real
kernels) and on the CPU with the kernel entry points replaced by their oracle formulas (checks the composition: argument
order, weights, attribute names, gradient flow); a third check pins the table below to the reference classes themselves."""
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
from conftest import assert_close

LAM = dict(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.25, lam_lp_op=2.0, lam_adv_op=0.03)
CASE = dict(seed=91, b=2, h=16, w=20, hd=9, wd=11)
LEAVES = ("rgb_out", "op_out", "d_gen", "latent")

# class name -> (attribute of the total, [(attribute, weight name, component key)] in the reference's order, argument keys)
TABLE = {
    "rgb_Loss": ("g_loss", [("g_adv_loss", "lam_adv", "adv"), ("g_gd_loss", "lam_gdl", "gd"), ("g_flow_loss", "lam_flow", "flow"),
                            ("g_int_loss", "lam_lp", "int")],
                 ("flow_pred", "flow_gt", "rgb_out", "rgb_tgt", "d_gen")),
    "rgb_vq_Loss": ("g_loss", [("g_adv_loss", "lam_adv", "adv"), ("g_gd_loss", "lam_gdl", "gd"), ("g_flow_loss", "lam_flow", "flow"),
                               ("g_int_loss", "lam_lp", "int"), ("g_latent_loss", "lam_latent", "lat")],
                    ("flow_pred", "flow_gt", "rgb_out", "rgb_tgt", "latent", "d_gen")),
    "op_loss": ("g_loss_op", [("g_int_loss_op", "lam_lp_op", "int_op"), ("g_adv_loss_op", "lam_adv_op", "adv")],
                ("op_out", "op_tgt", "d_gen")),
    "op_vq_Loss": ("g_loss_op", [("g_int_loss_op", "lam_lp_op", "int_op"), ("g_adv_loss_op", "lam_adv_op", "adv"),
                                 ("g_latent_loss", "lam_latent", "lat")],
                   ("op_out", "op_tgt", "d_gen", "latent")),
    "op_loss_v1": ("g_loss_op", [("g_int_loss_op", "lam_lp_op", "int_op")], ("op_out", "op_tgt")),
    "op_vq_Loss_v1": ("g_loss_op", [("g_int_loss_op", "lam_lp_op", "int_op"), ("g_latent_loss_op", "lam_latent", "lat")],
                      ("op_out", "op_tgt", "latent")),
    "Twostream_Loss": ("g_loss", [("g_adv_loss", "lam_adv", "adv"), ("g_gd_loss", "lam_gdl", "gd"), ("g_flow_loss", "lam_flow", "flow"),
                                  ("g_int_loss", "lam_lp", "int"), ("g_int_loss_op", "lam_lp_op", "int_op")],
                       ("flow_pred", "flow_gt", "rgb_out", "rgb_tgt", "op_out", "op_tgt", "d_gen")),
}


def inputs(device, dtype=torch.float32):
    t = {k: v.to(device=device, dtype=dtype) for k, v in synth.objective_inputs(CASE).items()}
    for k in LEAVES:
        t[k].requires_grad_(True)
    return t


def components(t):
    return dict(adv=O.adversarial_loss(t["d_gen"]), flow=O.flow_loss(t["flow_pred"], t["flow_gt"]),
                int=O.intensity_loss(t["rgb_out"], t["rgb_tgt"]), gd=O.gradient_loss(t["rgb_out"], t["rgb_tgt"]),
                int_op=O.intensity_loss(t["op_out"], t["op_tgt"]), lat=t["latent"].sum())


def expected(name, t):
    """(total, {attribute: value}) from the oracle components of `t`, weighted as TABLE says."""
    total_attr, terms, _ = TABLE[name]
    c = components(t)
    total = sum(LAM[w] * c[key] for _, w, key in terms)
    attrs = {a: c[key] for a, _, key in terms}
    attrs[total_attr] = total
    return total, attrs


def check_composed(device, rtol):
    for name, (total_attr, terms, argkeys) in TABLE.items():
        t = inputs(device)
        fn = getattr(A, name)(**LAM)
        got = fn(*[t[k] for k in argkeys])
        got.sum().backward()
        r = inputs("cpu", torch.float64)
        want, attrs = expected(name, r)
        want.backward()
        assert_close(got.detach().cpu().reshape(()), want.detach().reshape(()), rtol, name + " total")
        for a, v in attrs.items():
            assert isinstance(getattr(fn, a), float), name + "." + a
            assert_close(torch.tensor(getattr(fn, a)), v.detach().reshape(()), rtol, name + "." + a)
        used = {k for k in argkeys if k in LEAVES}
        for k in LEAVES:
            if k in used:
                assert_close(t[k].grad.cpu(), r[k].grad, rtol, name + " grad " + k)
            else:
                assert t[k].grad is None, name + " touched " + k
