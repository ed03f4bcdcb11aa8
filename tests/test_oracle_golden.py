"""CPU: the oracle restatement reproduces every fixture the live reference generated (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

import ammc_oracle as O
from ammcnet_aaai2021_b200 import synth
from conftest import load_golden, assert_close

MEM = ["mem_shipped", "mem_cfg1", "mem_k3", "mem_k1"]


def _mem_inputs(c):
    p = synth.memory_params(c["seed"], c["C"], c["D"], c["M"], c["k"])
    x = synth.features(c["seed"] + 1000, c["b"], c["C"], c["h"], c["w"])
    return p, x


@pytest.mark.parametrize("name", MEM)
def test_memory_forward(name):
    c, g = load_golden(name)
    p, x = _mem_inputs(c)
    o = O.memory_module_forward(x, p["enc.weight"], p["enc.bias"], p["quantize.embed"], p["dec.weight"],
                                p["dec.bias"], c["k"])
    assert np.array_equal(o["idx_topk"].numpy(), g["idx_topk"])
    assert np.array_equal(o["quantize_topk"].numpy(), g["read"])
    assert_close(o["out"], g["out"], 1e-6, name + ".out")
    assert_close(o["diff"], g["diff"], 1e-6, name + ".diff")
    assert_close(o["quantize"], g["q1"], 1e-6, name + ".q1")
    assert_close(o["sse_per_frame"], g["sse_per_frame"], 1e-6, name + ".sse")


@pytest.mark.parametrize("name", MEM)
def test_memory_training_ema(name):
    c, g = load_golden(name)
    p, x = _mem_inputs(c)
    x2 = synth.features(c["seed"] + 2000, c["b"], c["C"], c["h"], c["w"])
    cs, ea, em = p["quantize.cluster_size"], p["quantize.embed_avg"], p["quantize.embed"]
    for step, xs in enumerate((x, x2)):
        o = O.memory_module_forward(xs, p["enc.weight"], p["enc.bias"], em, p["dec.weight"], p["dec.bias"], c["k"],
                                    training=True, cluster_size=cs, embed_avg=ea)
        cs, ea, em = o["cluster_size"], o["embed_avg"], o["embed"]
        assert_close(o["out"], g[f"train{step}_out"], 1e-5, f"{name}.train{step}.out")
        assert_close(cs, g[f"train{step}_cluster_size"], 1e-5, f"{name}.train{step}.cs")
        assert_close(ea, g[f"train{step}_embed_avg"], 1e-5, f"{name}.train{step}.avg")
        assert_close(em, g[f"train{step}_embed"], 1e-4, f"{name}.train{step}.embed")


@pytest.mark.parametrize("name", MEM)
def test_memory_backward(name):
    c, g = load_golden(name)
    p, x = _mem_inputs(c)
    o = O.memory_module_forward(x, p["enc.weight"], p["enc.bias"], p["quantize.embed"], p["dec.weight"],
                                p["dec.bias"], c["k"])
    gen = torch.Generator().manual_seed(c["seed"] + 3000)
    r_out = torch.randn(o["out"].shape, generator=gen)
    r_q1 = torch.randn(o["quantize"].shape, generator=gen)
    ob = O.memory_module_backward(x, p["enc.weight"], p["enc.bias"], p["quantize.embed"], p["dec.weight"],
                                  o["idx_topk"], o["z"], r_out, torch.tensor(3.0), r_q1)
    for k in ("gx", "g_enc_w", "g_enc_b", "g_dec_w", "g_dec_b"):
        assert_close(ob[k], g[k], 1e-4, f"{name}.{k}")


@pytest.mark.parametrize("name", ["amft_c64", "amft_c512"])
def test_amft(name):
    c, g = load_golden(name)
    p = synth.amft_params(c["seed"], c["C"])
    zx = synth.features(c["seed"] + 1000, c["b"], c["C"], c["h"], c["w"])
    zy = synth.features(c["seed"] + 2000, c["b"], c["C"], c["h"], c["w"])
    x, y, _ = O.amft_forward(zx, zy, p)
    assert_close(x, g["x"], 1e-5, name + ".x")
    assert_close(y, g["y"], 1e-5, name + ".y")
    tx, ty, stats = O.amft_forward(zx, zy, p, training=True)
    assert_close(tx, g["train_x"], 1e-4, name + ".train_x")
    assert_close(ty, g["train_y"], 1e-4, name + ".train_y")
    for k, v in stats.items():
        assert_close(v, g["stat_" + k], 1e-5, name + "." + k)


def test_psnr():
    c, g = load_golden("psnr")
    gen, gt = synth.frames(c["seed"], c["b"], c["c"], c["h"], c["w"])
    assert_close(O.psnr_per_frame(gen, gt), g["per_frame"], 1e-6, "psnr.per_frame")
    assert_close(O.psnr_error(gen, gt), g["batch_mean"], 1e-6, "psnr.batch")


@pytest.mark.parametrize("ds", ["ped2", "avenue", "shanghaitech"])
def test_score_reduction_known_answers(ds):
    """Recorded per-frame score pickles of the reference's real runs -> scores (bit-exact) and the KAT sums of BASELINE.md."""
    c, g = load_golden("scores_" + ds)
    offs = np.concatenate([[0], np.cumsum(g["lengths"])])
    img = [g["img"][offs[i]:offs[i + 1]] for i in range(len(g["lengths"]))]
    fea = [g["fea"][offs[i]:offs[i + 1]] for i in range(len(g["lengths"]))]
    s = O.score_reduce(img, fea, tuple(c["lam"]))
    assert s.dtype == np.float32 and np.array_equal(s, g["scores"])
    kat = {"ped2": (1962, 910.9196372), "avenue": (15240, 11922.8437268), "shanghaitech": (40363, 28164.4554798)}[ds]
    assert len(s) == kat[0]
    assert abs(float(np.sum(s.astype(np.float64))) - kat[1]) < 1e-4
    assert abs(O.roc_auc(g["labels_kept"] if "labels_kept" in g else
                         np.concatenate([g["labels"][offs[i] + 4:offs[i + 1]] for i in range(len(g["lengths"]))]),
                         s) - float(g["auc"])) < 1e-12


def test_record_assembly():
    """Record layout of the reference scoring loop, from records the reference itself wrote (gen_golden.gen_records)."""
    c, g = load_golden("records")
    for v, T in enumerate(c["lengths"]):
        rgb_img = g[f"rgb_img_pred_records_{v}"]
        rgb_fea = g[f"rgb_fea_comm_records_{v}"]
        n_clips = T - 4
        psnr_clip = rgb_img[4:]
        groups = np.array([rgb_fea[4 + 16 * i] for i in range((n_clips + 15) // 16)], np.float32)
        img, fea = O.assemble_video_records(psnr_clip, groups, clip_len=5)
        assert np.array_equal(img, rgb_img) and np.array_equal(fea, rgb_fea)
        op_fea = g[f"op_fea_comm_records_{v}"]
        op_img = g[f"op_img_pred_records_{v}"]
        ogroups = np.array([op_fea[3 + 16 * i] for i in range((n_clips + 15) // 16)], np.float32)
        img2, fea2 = O.assemble_video_records(op_img[3:3 + n_clips], ogroups, clip_len=4, tail_copy=True)
        assert np.array_equal(fea2, op_fea) and np.array_equal(img2, op_img)


@pytest.mark.parametrize("name", ["gen_64", "gen_96x160"])
def test_generator_oracle_vs_reference_golden(name):
    """The oracle's functional restatement of twostream.forward (unet.py:981-1007) against the live reference's outputs;
    also pins synth.generator_params to the reference's state_dict layout (222 entries, loaded strict by gen_golden)."""
    import torch
    from ammcnet_aaai2021_b200 import synth, host_model
    c, g = load_golden(name)
    p = synth.generator_params(c["seed"])
    assert set(p) == set(host_model.get_twostream().state_dict()) and len(p) == 222
    rgb, op = synth.generator_inputs(c["seed"] + 500, c["b"], c["h"], c["w"])
    with torch.no_grad():
        ry, oy, (rd, od), (rq, oq) = O.twostream_forward(rgb, op, p, 2)
    for a, key in ((ry, "rgb_y"), (oy, "op_y"), (rd, "rgb_diff"), (od, "op_diff"), (rq, "rgb_q1"), (oq, "op_q1")):
        assert_close(a, g[key], 1e-5, key)


def test_preprocess_oracle_vs_reference_loaders():
    """cv2.resize restatement + the loaders' arithmetic (two_stream_dataset.py:72-99) against what the reference's own
    _load_frame / _load_op returned for the same decoded inputs: bit-exact, arrays for small cases, digests for the
    datasets' native sizes."""
    import hashlib
    c, g = load_golden("preprocess")
    n = 0
    for name, bgr, flow, size in synth.preprocess_inputs(c["seed"]):
        rgb, op = O.preprocess_frame(bgr, size), O.preprocess_flow(flow, size)
        meta = c["cases"][name]
        if "rgb_sha256" in meta:
            assert hashlib.sha256(rgb.tobytes()).hexdigest() == meta["rgb_sha256"], name
            assert hashlib.sha256(op.tobytes()).hexdigest() == meta["op_sha256"], name
        else:
            assert np.array_equal(bgr, g[name + "_bgr"]) and np.array_equal(flow, g[name + "_flow"]), name
            assert np.array_equal(rgb, g[name + "_rgb_out"]), name
            assert np.array_equal(op, g[name + "_op_out"]), name
        assert rgb.dtype == np.float32 and rgb.shape == (3, size[1], size[0]) and op.shape == (2, size[1], size[0])
        n += 1
    assert n == len(synth.PREPROCESS_CASES) == 8


def test_frame_losses_oracle_vs_reference():
    """Intensity_Loss / Gradient_Loss restatement (losses_utils.py:17-59) against the reference's values and autograd."""
    c, g = load_golden("losses")
    for name, cs in c["cases"].items():
        gen, gt = synth.frames(cs["seed"], cs["b"], cs["C"], cs["h"], cs["w"])
        gen.requires_grad_(True)
        li, lg = O.intensity_loss(gen, gt), O.gradient_loss(gen, gt)
        assert_close(li.detach(), g[name + "_int"], 1e-6, name + ".int")
        assert_close(lg.detach(), g[name + "_gd"], 1e-6, name + ".gd")
        gi = torch.autograd.grad(li, gen, retain_graph=True)[0]
        gg = torch.autograd.grad(lg, gen)[0]
        if name + "_g_int" in g:
            assert_close(gi, g[name + "_g_int"], 1e-6, name + ".g_int")
            assert_close(gg, g[name + "_g_gd"], 1e-6, name + ".g_gd")
        else:
            assert_close((gi.double() * gt.double()).sum(), g[name + "_g_int_sum"], 1e-6, name + ".g_int_sum")
            assert_close((gg.double() * gt.double()).sum(), g[name + "_g_gd_sum"], 1e-6, name + ".g_gd_sum")


def test_training_objectives_oracle_vs_reference():
    """Flow_Loss / Adversarial_Loss / Discriminate_Loss / Twostream_vq_Loss restatement (losses_utils.py:10-15,103-113,
    loss_zoo.py:307-350) against the values, stored attributes and autograd gradients of the reference itself."""
    c, g = load_golden("objectives")
    lam = c["lambdas"]
    for name, cs in c["cases"].items():
        t = synth.objective_inputs(cs)
        assert_close(O.flow_loss(t["flow_pred"], t["flow_gt"]), g[name + "_flow"], 1e-6, name + ".flow")
        assert_close(O.adversarial_loss(t["d_gen"]), g[name + "_adv"], 1e-6, name + ".adv")
        dr, df = t["d_real"].clone().requires_grad_(True), t["d_gen"].clone().requires_grad_(True)
        ld = O.discriminate_loss(dr, df)
        assert_close(ld.detach(), g[name + "_dis"], 1e-6, name + ".dis")
        g_dr, g_df = torch.autograd.grad(ld, (dr, df))
        assert_close(g_dr, g[name + "_g_dis_real"], 1e-6, name + ".g_dis_real")
        assert_close(g_df, g[name + "_g_dis_fake"], 1e-6, name + ".g_dis_fake")
        leaves = {k: t[k].clone().requires_grad_(True) for k in ("rgb_out", "op_out", "latent", "d_gen")}
        loss, parts = O.twostream_vq_loss(lam, t["flow_pred"], t["flow_gt"], leaves["rgb_out"], t["rgb_tgt"], leaves["op_out"],
                                          t["op_tgt"], leaves["latent"], leaves["d_gen"])
        assert_close(loss.detach(), g[name + "_g_loss"], 1e-6, name + ".g_loss")
        for k, v in parts.items():
            assert_close(v.detach().reshape(()), g[name + "_attr_" + k], 1e-6, name + "." + k)
        grads = torch.autograd.grad(loss, list(leaves.values()))
        for (k, leaf), gr in zip(leaves.items(), grads):
            if name + "_dg_" + k in g:
                assert_close(gr, g[name + "_dg_" + k], 1e-6, name + ".dg_" + k)
            else:
                assert_close((gr.double() * t[k].double()).sum(), g[name + "_dg_" + k + "_sum"], 1e-6, name + ".dg_" + k + "_sum")
