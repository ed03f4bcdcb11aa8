"""Generate tests/golden/*.npz by running the UNMODIFIED reference (needs /root/reference; build container only).

    python oracle/gen_golden.py            # rewrites every fixture and checks the oracle against each

Inputs and weights are NOT stored: they are regenerated from seeds by `ammcnet_aaai2021_b200.synth`
(CPU torch generators are reproducible for a fixed torch build, and the GPU box runs this same image).
Each fixture holds the case description (`meta` json) and the reference's outputs.  The script fails if
the oracle restatement (`oracle/ammc_oracle.py`) disagrees with the live reference.
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ammc_oracle as O                      # noqa: E402
import ref_harness                           # noqa: E402
import digest                                # noqa: E402
from ammcnet_aaai2021_b200 import synth     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# ---- case tables (shared with tests/cases.py through the fixture `meta`) --------------------------
MEM_CASES = {
    # name: b, C, h, w, D, M, k, seed
    "mem_shipped": dict(b=2, C=512, h=8, w=16, D=64, M=256, k=2, seed=11),   # shipped D/M/k (net_params/*.pkl)
    "mem_cfg1":    dict(b=1, C=512, h=8, w=8, D=512, M=10, k=2, seed=12),    # BASELINE.json configs[0]
    "mem_k3":      dict(b=3, C=96, h=4, w=8, D=32, M=50, k=3, seed=13),      # ragged, k=3
    "mem_k1":      dict(b=1, C=64, h=5, w=7, D=16, M=7, k=1, seed=14),       # odd sizes, k=1
}
AMFT_CASES = {
    "amft_c64":  dict(b=2, C=64, h=8, w=8, seed=21),
    "amft_c512": dict(b=1, C=512, h=8, w=32, seed=22),
}


def _save(name, meta, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), meta=json.dumps(meta), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items()})


def _close(a, b, tol, what):
    a = torch.as_tensor(np.asarray(a.detach() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach() if torch.is_tensor(b) else b), dtype=torch.float64)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = max(b.abs().max().item(), 1e-30) if b.numel() else 1.0
    assert err <= tol * scale, f"oracle != reference for {what}: max err {err} (scale {scale})"


def build_ref_memory(ref_unet, c, params, res=True):
    m = (ref_unet.enc_quan_dec_res_topk if res else ref_unet.enc_quan_dec_topk)(c["C"], c["D"], c["M"], k=c["k"])
    sd = {("quan." if res else "") + k: v.clone() for k, v in params.items()}
    m.load_state_dict(sd, strict=True)
    return m


def gen_memory(ref_unet):
    for name, c in MEM_CASES.items():
        p = synth.memory_params(c["seed"], c["C"], c["D"], c["M"], c["k"])
        x = synth.features(c["seed"] + 1000, c["b"], c["C"], c["h"], c["w"])
        # ---- eval forward -----------------------------------------------------------------------
        m = build_ref_memory(ref_unet, c, p).eval()
        with torch.no_grad():
            out, diff, q1 = m(x.clone())
            z = m.quan.enc(x).permute(0, 2, 3, 1)
            read, diff_q, _ = m.quan.quantize(z)
        o = O.memory_module_forward(x, p["enc.weight"], p["enc.bias"], p["quantize.embed"],
                                    p["dec.weight"], p["dec.bias"], c["k"])
        _close(o["out"], out, 1e-6, name + ".out")
        _close(o["diff"], diff, 1e-6, name + ".diff")
        _close(o["quantize"], q1, 0, name + ".q1")
        _close(o["quantize_topk"], read, 0, name + ".read")
        # margins of the reference's ranking (for the no-tie filter in the GPU tests)
        dsort = o["dist"].sort(1)[0]
        kk = min(c["k"] + 1, c["M"])
        # ---- training forward x2 (EMA), reference mutates its buffers in place -------------------
        mt = build_ref_memory(ref_unet, c, p).train()
        x2 = synth.features(c["seed"] + 2000, c["b"], c["C"], c["h"], c["w"])
        tr = {}
        cs, ea, em = p["quantize.cluster_size"], p["quantize.embed_avg"], p["quantize.embed"]
        for step, xs in enumerate((x, x2)):
            with torch.no_grad():
                out_t, diff_t, _ = mt(xs.clone())
            ot = O.memory_module_forward(xs, p["enc.weight"], p["enc.bias"], em, p["dec.weight"], p["dec.bias"],
                                         c["k"], training=True, cluster_size=cs, embed_avg=ea)
            cs, ea, em = ot["cluster_size"], ot["embed_avg"], ot["embed"]
            sdt = mt.state_dict()
            _close(ot["out"], out_t, 1e-6, f"{name}.train{step}.out")
            _close(cs, sdt["quan.quantize.cluster_size"], 1e-6, f"{name}.train{step}.cluster_size")
            _close(ea, sdt["quan.quantize.embed_avg"], 1e-6, f"{name}.train{step}.embed_avg")
            _close(em, sdt["quan.quantize.embed"], 1e-5, f"{name}.train{step}.embed")
            tr[f"train{step}_out"] = out_t
            tr[f"train{step}_diff"] = diff_t
            tr[f"train{step}_cluster_size"] = sdt["quan.quantize.cluster_size"].clone()
            tr[f"train{step}_embed_avg"] = sdt["quan.quantize.embed_avg"].clone()
            tr[f"train{step}_embed"] = sdt["quan.quantize.embed"].clone()
        # ---- backward through the reference with autograd ---------------------------------------
        mb = build_ref_memory(ref_unet, c, p).eval()
        xg = x.clone().requires_grad_(True)
        g = torch.Generator().manual_seed(c["seed"] + 3000)
        out_b, diff_b, q1_b = mb(xg)
        r_out = torch.randn(out_b.shape, generator=g)
        r_q1 = torch.randn(q1_b.shape, generator=g)
        g_diff = 3.0
        (out_b * r_out).sum().add(g_diff * diff_b.sum()).add((q1_b * r_q1).sum()).backward()
        ob = O.memory_module_backward(x, p["enc.weight"], p["enc.bias"], p["quantize.embed"], p["dec.weight"],
                                      o["idx_topk"], o["z"], r_out, torch.tensor(g_diff), r_q1)
        _close(ob["gx"], xg.grad, 2e-5, name + ".gx")
        _close(ob["g_enc_w"], mb.quan.enc.weight.grad, 2e-5, name + ".g_enc_w")
        _close(ob["g_enc_b"], mb.quan.enc.bias.grad, 2e-5, name + ".g_enc_b")
        _close(ob["g_dec_w"], mb.quan.dec.weight.grad, 2e-5, name + ".g_dec_w")
        _close(ob["g_dec_b"], mb.quan.dec.bias.grad, 2e-5, name + ".g_dec_b")
        _save(name, dict(kind="memory", **c), out=out, diff=diff, q1=q1, read=read,
              idx_topk=o["idx_topk"].to(torch.int32), dist_sorted=dsort[:, :kk],
              sse_per_frame=o["sse_per_frame"],
              gx=xg.grad, g_enc_w=mb.quan.enc.weight.grad, g_enc_b=mb.quan.enc.bias.grad,
              g_dec_w=mb.quan.dec.weight.grad, g_dec_b=mb.quan.dec.bias.grad, **tr)


def gen_amft(ref_unet):
    for name, c in AMFT_CASES.items():
        p = synth.amft_params(c["seed"], c["C"])
        zx = synth.features(c["seed"] + 1000, c["b"], c["C"], c["h"], c["w"])
        zy = synth.features(c["seed"] + 2000, c["b"], c["C"], c["h"], c["w"])
        m = ref_unet.bridge(in_c=c["C"])
        m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
        m.eval()
        with torch.no_grad():
            yx, yy = m(zx, zy)
        ox, oy, _ = O.amft_forward(zx, zy, p)
        _close(ox, yx, 1e-6, name + ".x")
        _close(oy, yy, 1e-6, name + ".y")
        # training-mode forward + backward (batch statistics)
        mt = ref_unet.bridge(in_c=c["C"])
        mt.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
        mt.train()
        zxg, zyg = zx.clone().requires_grad_(True), zy.clone().requires_grad_(True)
        tx, ty = mt(zxg, zyg)
        g = torch.Generator().manual_seed(c["seed"] + 3000)
        rx, ry = torch.randn(tx.shape, generator=g), torch.randn(ty.shape, generator=g)
        ((tx * rx).sum() + (ty * ry).sum()).backward()
        otx, oty, stats = O.amft_forward(zx, zy, p, training=True)
        _close(otx, tx, 1e-5, name + ".train.x")
        _close(oty, ty, 1e-5, name + ".train.y")
        sdt = mt.state_dict()
        for kname, v in stats.items():
            _close(v, sdt[kname], 1e-5, name + ".train." + kname)
        extra = {}
        for i, (pn, pv) in enumerate(mt.named_parameters()):
            if c["C"] <= 64 or pv.grad.dim() == 1:      # full gradients: small case, and every BatchNorm vector
                extra["g_" + pn] = pv.grad
            else:                                       # shipped C=512 conv weights: reduced forms (oracle/digest.py)
                for dk, dv in digest.weight_grad_digest(pv.grad, c["seed"] + 4000 + i).items():
                    extra["gd_%s_%s" % (dk, pn)] = dv
        _save(name, dict(kind="amft", **c), x=yx, y=yy, train_x=tx, train_y=ty,
              g_zx=zxg.grad, g_zy=zyg.grad,
              **{"stat_" + k: sdt[k] for k in stats}, **extra)


def gen_psnr(ref_utils):
    gen, gt = synth.frames(31, 5, 3, 64, 48)
    per = torch.stack([ref_utils.psnr_error(gen[i:i + 1], gt[i:i + 1]) for i in range(gen.shape[0])])
    batch = ref_utils.psnr_error(gen, gt)
    _close(O.psnr_per_frame(gen, gt), per, 1e-6, "psnr.per_frame")
    _close(O.psnr_error(gen, gt), batch, 1e-6, "psnr.batch")
    _save("psnr", dict(kind="psnr", seed=31, b=5, c=3, h=64, w=48), per_frame=per, batch_mean=batch)


LAM = {"avenue": (0.04, 0.65), "ped2": (0.01, 0.55), "shanghaitech": (0.13, 0.60)}   # test_helper.py:565-569


def gen_scores(ref_eval):
    """Run the reference's own img_pred_fea_comm_single_auc on its recorded pickles; capture the score vector
    it hands to sklearn (labels are synthetic: the GT .mat/.npy files live in the datasets)."""
    from sklearn import metrics as skm
    for ds, lam in LAM.items():
        pk = os.path.join(ref_harness.REFERENCE_ROOT, "Code", "ammcnet_os", "model_result_save", ds,
                          "img_pred_fea_comm_rgb_auc", "save_pickle", ds)
        rec = pickle.load(open(pk, "rb"))
        img = [np.array(a, dtype=np.float32) for a in rec["rgb_img_pred_records"]]
        fea = [np.array(a, dtype=np.float32) for a in rec["rgb_fea_comm_records"]]
        rng = np.random.RandomState(7)
        labels = [(rng.rand(len(a)) < 0.3).astype(np.int8) for a in img]
        captured = {}
        orig_call = ref_eval.GroundTruthLoader.__call__
        orig_roc = ref_eval.metrics.roc_curve

        def fake_roc(lbl, scores, pos_label=None):
            captured["scores"] = np.asarray(scores, dtype=np.float32)
            captured["labels"] = np.asarray(lbl)
            return orig_roc(lbl, scores, pos_label=pos_label)

        ref_eval.GroundTruthLoader.__call__ = lambda self, dataset=None: labels
        ref_eval.metrics.roc_curve = fake_roc
        try:
            ret = ref_eval.evaluate("img_pred_fea_comm_rgb_auc", pk, lam)
        finally:
            ref_eval.GroundTruthLoader.__call__ = orig_call
            ref_eval.metrics.roc_curve = orig_roc
        s = O.score_reduce(img, fea, lam)
        assert np.array_equal(s, captured["scores"]), ds
        fpr, tpr, _ = skm.roc_curve(captured["labels"], captured["scores"], pos_label=0)
        auc = skm.auc(fpr, tpr)
        assert abs(O.roc_auc(captured["labels"], captured["scores"]) - auc) < 1e-12
        assert round(auc, 3) == ret["auc"]
        print(ds, "T=", len(s), "first4=", s[:4], "sum=", float(np.sum(s.astype(np.float64))), "auc(synth labels)=", auc)
        lens = np.array([len(a) for a in img], np.int64)
        _save("scores_" + ds, dict(kind="scores", dataset=ds, lam=lam),
              lengths=lens, img=np.concatenate(img), fea=np.concatenate(fea),
              labels=np.concatenate(labels), scores=captured["scores"], auc=np.float64(auc))


def gen_records():
    """Drive the reference inference/scoring loop (test_helper.py:387-488) with a stand-in generator and
    synthetic clip datasets; store the pickled records it writes."""
    import types
    import tempfile
    import logging
    stub = types.ModuleType("Code.main.constant_test")
    stub.const = types.SimpleNamespace(gpu_idx="0")
    sys.modules["Code.main.constant_test"] = stub
    torch.Tensor.cuda = lambda s, *a, **k: s
    torch.nn.Module.cuda = lambda s, *a, **k: s
    import Code.run_helper.test_helper as TH
    real_loader = TH.DataLoader
    TH.DataLoader = lambda ds, batch_size, shuffle, num_workers: real_loader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=0)

    lengths = [23, 40, 9]
    H = W = 16

    class Clips(torch.utils.data.Dataset):
        def __init__(self, folder, clip_len, data_type):
            vid = int(os.path.basename(folder))
            self.T, self.clip, self.ch = lengths[vid], clip_len, (3 if data_type == "rgb" else 2)
            g = torch.Generator().manual_seed(100 + vid * 2 + (data_type == "op"))
            n = self.T if data_type == "rgb" else self.T - 1
            self.fr = torch.rand((n, self.ch, H, W), generator=g) * 2 - 1

        def __len__(self):
            return self.fr.shape[0] - self.clip + 1

        def __getitem__(self, i):
            return self.fr[i:i + self.clip]

    class FakeGen(torch.nn.Module):
        def forward(self, rgb_in, op_in):
            rgb = (0.9 * rgb_in[:, -3:] + 0.05 * rgb_in[:, :3]).clamp(-1, 1)
            op = op_in[:, -2:] * 0.8
            return rgb, op, (rgb_in.pow(2).mean().unsqueeze(0), op_in.pow(2).mean().unsqueeze(0)), None

    with tempfile.TemporaryDirectory() as td:
        for root in ("rgb", "op"):
            for v in range(len(lengths)):
                os.makedirs(os.path.join(td, root, "%02d" % v))
        pk = os.path.join(td, "out.pkl")
        TH.gen_loss_file_twostream_normal_all(FakeGen().eval(), Clips, "unet_vq_twostream",
                                              (os.path.join(td, "rgb"), os.path.join(td, "op")), (5, 4),
                                              "synthetic", pk, "psnr", logging.getLogger("gen_golden"))
        rec = pickle.load(open(pk, "rb"))
    _save("records", dict(kind="records", lengths=lengths, H=H, W=W),
          **{f"{key}_{v}": rec[key][v] for key in ("rgb_img_pred_records", "rgb_fea_comm_records",
                                                   "op_img_pred_records", "op_fea_comm_records")
             for v in range(len(lengths))})


GEN_CASES = {
    # the shipped twostream generator (12/6 -> 3/2 channels, D=64, M=256, k=2) on small frames
    "gen_64": dict(b=2, h=64, w=64, seed=31),
    "gen_96x160": dict(b=1, h=96, w=160, seed=32),      # non-square, 160 = one full + one partial 128-pixel segment
}


def gen_generator(ref_unet):
    """Whole twostream generator, eval mode: reference vs the oracle's functional restatement (unet.py:981-1007)."""
    for name, c in GEN_CASES.items():
        p = synth.generator_params(c["seed"])
        ref = ref_unet.twostream(12, 3, 6, 2, embed_dim=64, n_embed=256, k=2)
        ref.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
        ref.eval()
        rgb, op = synth.generator_inputs(c["seed"] + 500, c["b"], c["h"], c["w"])
        with torch.no_grad():
            ry, oy, (rd, od), (rq, oq) = ref(rgb, op)
            ary, aoy, (ard, aod), (arq, aoq) = O.twostream_forward(rgb, op, p, 2)
        for a, b_, what in ((ary, ry, "rgb_y"), (aoy, oy, "op_y"), (ard, rd, "rgb_diff"), (aod, od, "op_diff"),
                            (arq, rq, "rgb_q1"), (aoq, oq, "op_q1")):
            _close(a, b_, 1e-6, f"{name}.{what}")
        _save(name, dict(kind="generator", **c), rgb_y=ry, op_y=oy, rgb_diff=rd, op_diff=od, rgb_q1=rq, op_q1=oq)


def gen_preprocess():
    """Frame / flow loaders of the reference (two_stream_dataset._load_frame / _load_op) on synthetic decoded inputs.
    The JPEG decoder is replaced by the synthetic BGR array (decoding is not part of the restated path); .flo files are
    written to a temp dir and read back by the reference's own readFlow."""
    import hashlib
    import tempfile
    import torchvision.transforms as T
    import cv2
    ref_harness.import_reference()            # installs the stub modules and puts the reference on sys.path
    import Code.dataset.two_stream_dataset as ds
    tf_rgb = T.Compose([T.ToTensor(), T.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])      # two_stream_dataset.py:501-505
    out, meta = {}, {}
    tmp = tempfile.mkdtemp()
    for name, bgr, flow, (W, H) in synth.preprocess_inputs(20200525):
        h0, w0 = bgr.shape[:2]
        ds.img_decode = lambda path, _a=bgr: _a.copy()
        ref_rgb = ds._load_frame("synthetic.jpg", img_size=(W, H), transform=tf_rgb).numpy()
        flo = os.path.join(tmp, name + ".flo")
        with open(flo, "wb") as f:
            np.array([202021.25], np.float32).tofile(f)
            np.array([w0, h0], np.int32).tofile(f)
            flow.tofile(f)
        ref_op = ds._load_op(flo, img_size=(W, H)).numpy()
        mine_rgb, mine_op = O.preprocess_frame(bgr, (W, H)), O.preprocess_flow(flow, (W, H))
        assert np.array_equal(mine_rgb, ref_rgb), f"oracle != reference _load_frame for {name}"
        assert np.array_equal(mine_op, ref_op), f"oracle != reference _load_op for {name}"
        assert np.array_equal(O.resize_linear_u8(bgr, W, H), cv2.resize(bgr, (W, H)))
        assert np.array_equal(O.resize_linear_f32(flow, W, H), cv2.resize(flow, (W, H)))
        meta[name] = dict(src=[h0, w0], dst=[W, H])
        if h0 * w0 <= 20000:                      # small cases: the arrays themselves
            out[name + "_bgr"], out[name + "_flow"] = bgr, flow
            out[name + "_rgb_out"], out[name + "_op_out"] = ref_rgb, ref_op
        else:                                     # dataset sizes: inputs come from the seed, outputs are pinned by digest
            meta[name]["rgb_sha256"] = hashlib.sha256(ref_rgb.tobytes()).hexdigest()
            meta[name]["op_sha256"] = hashlib.sha256(ref_op.tobytes()).hexdigest()
    _save("preprocess", dict(kind="preprocess", seed=20200525, cases=meta, cv2=cv2.__version__), **out)


LOSS_CASES = {"loss_rgb_small": dict(b=2, C=3, h=16, w=20, seed=41), "loss_op_small": dict(b=3, C=2, h=9, w=7, seed=42),
              "loss_rgb_256": dict(b=2, C=3, h=256, w=256, seed=43)}


def gen_losses():
    """Intensity_Loss / Gradient_Loss of the reference (losses_utils.py:17-59) with autograd gradients.  The module reads
    the training configuration at import and moves its filters with .cuda(): a stub `const` and an identity Tensor.cuda
    stand in for both here (the reference file itself is imported unmodified)."""
    import types
    ref_harness.import_reference()
    stub = types.ModuleType("Code.main.constant_train")
    stub.const = types.SimpleNamespace(gpu_idx="0")
    sys.modules["Code.main.constant_train"] = stub
    import Code.models.losses.losses_utils as LU
    cuda_orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = {}
        for name, c in LOSS_CASES.items():
            gen_f, gt_f = synth.frames(c["seed"], c["b"], c["C"], c["h"], c["w"])
            gen_f.requires_grad_(True)
            li = LU.Intensity_Loss()(gen_f, gt_f)
            lg = LU.Gradient_Loss(channels=c["C"])(gen_f, gt_f)
            (gi,) = torch.autograd.grad(li, gen_f, retain_graph=True)
            (gg,) = torch.autograd.grad(lg, gen_f)
            g2 = gen_f.detach().clone().requires_grad_(True)
            oi, og = O.intensity_loss(g2, gt_f), O.gradient_loss(g2, gt_f)
            _close(oi, li, 1e-6, name + ".intensity")
            _close(og, lg, 1e-6, name + ".gradient")
            _close(torch.autograd.grad(oi, g2, retain_graph=True)[0], gi, 1e-6, name + ".d_intensity")
            _close(torch.autograd.grad(og, g2)[0], gg, 1e-6, name + ".d_gradient")
            out[name + "_int"], out[name + "_gd"] = li.detach(), lg.detach()
            if c["h"] * c["w"] <= 1024:
                out[name + "_g_int"], out[name + "_g_gd"] = gi, gg
            else:                                                       # large case: gradients pinned by two projections
                out[name + "_g_int_sum"] = (gi.double() * gt_f.double()).sum()
                out[name + "_g_gd_sum"] = (gg.double() * gt_f.double()).sum()
    finally:
        torch.Tensor.cuda = cuda_orig
    _save("losses", dict(kind="losses", cases=LOSS_CASES), **out)


OBJECTIVE_CASES = {"obj_small": dict(b=2, h=16, w=20, hd=9, wd=11, seed=61), "obj_odd": dict(b=3, h=9, w=7, hd=6, wd=5, seed=62),
                   "obj_256": dict(b=2, h=256, w=256, hd=33, wd=33, seed=63)}
OBJECTIVE_LAMBDAS = dict(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.25, lam_lp_op=2.0)


def gen_objectives():
    """Flow_Loss / Adversarial_Loss / Discriminate_Loss (losses_utils.py:10-15,103-113) and Twostream_vq_Loss
    (loss_zoo.py:307-350) of the reference, imported unmodified, with autograd gradients; stubs as in gen_losses."""
    import types
    ref_harness.import_reference()
    stub = types.ModuleType("Code.main.constant_train")
    stub.const = types.SimpleNamespace(gpu_idx="0")
    sys.modules["Code.main.constant_train"] = stub
    import Code.models.losses.losses_utils as LU
    import Code.models.losses.loss_zoo as LZ
    cuda_orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = {}
        for name, c in OBJECTIVE_CASES.items():
            t = synth.objective_inputs(c)
            small = c["h"] * c["w"] <= 1024
            # the three element-wise objectives on their own
            fp = t["flow_pred"].clone().requires_grad_(True)
            lf = LU.Flow_Loss()(fp, t["flow_gt"])
            (g_fp,) = torch.autograd.grad(lf, fp)
            dg = t["d_gen"].clone().requires_grad_(True)
            la = LU.Adversarial_Loss()(dg)
            (g_dg,) = torch.autograd.grad(la, dg)
            dr, df = t["d_real"].clone().requires_grad_(True), t["d_gen"].clone().requires_grad_(True)
            ld = LU.Discriminate_Loss()(dr, df)
            g_dr, g_df = torch.autograd.grad(ld, (dr, df))
            _close(O.flow_loss(t["flow_pred"], t["flow_gt"]), lf, 1e-6, name + ".flow")
            _close(O.adversarial_loss(t["d_gen"]), la, 1e-6, name + ".adv")
            _close(O.discriminate_loss(t["d_real"], t["d_gen"]), ld, 1e-6, name + ".dis")
            out[name + "_flow"], out[name + "_adv"], out[name + "_dis"] = lf.detach(), la.detach(), ld.detach()
            out[name + "_g_adv"], out[name + "_g_dis_real"], out[name + "_g_dis_fake"] = g_dg, g_dr, g_df
            if small:
                out[name + "_g_flow"] = g_fp
            else:
                out[name + "_g_flow_sum"] = (g_fp.double() * t["flow_gt"].double()).sum()
            # the generator objective: gradients w.r.t. everything the generator produces
            leaves = {k: t[k].clone().requires_grad_(True) for k in ("rgb_out", "op_out", "latent", "d_gen")}
            fn = LZ.Twostream_vq_Loss(**OBJECTIVE_LAMBDAS)
            g_loss = fn(t["flow_pred"], t["flow_gt"], leaves["rgb_out"], t["rgb_tgt"], leaves["op_out"], t["op_tgt"],
                        leaves["latent"], leaves["d_gen"])
            grads = torch.autograd.grad(g_loss, list(leaves.values()))
            o_loss, o_parts = O.twostream_vq_loss(OBJECTIVE_LAMBDAS, t["flow_pred"], t["flow_gt"], t["rgb_out"], t["rgb_tgt"],
                                                  t["op_out"], t["op_tgt"], t["latent"], t["d_gen"])
            _close(o_loss, g_loss, 1e-6, name + ".g_loss")
            for k in ("g_loss", "g_adv_loss", "g_flow_loss", "g_int_loss", "g_gd_loss", "g_int_loss_op", "g_latent_loss"):
                _close(o_parts[k], getattr(fn, k), 1e-6, name + "." + k)
                out[name + "_attr_" + k] = np.float64(getattr(fn, k))
            out[name + "_g_loss"] = g_loss.detach()
            for (k, leaf), g in zip(leaves.items(), grads):
                if small or k in ("latent", "d_gen"):
                    out[name + "_dg_" + k] = g
                else:
                    out[name + "_dg_" + k + "_sum"] = (g.double() * t[k].double()).sum()
    finally:
        torch.Tensor.cuda = cuda_orig
    _save("objectives", dict(kind="objectives", cases=OBJECTIVE_CASES, lambdas=OBJECTIVE_LAMBDAS), **out)


def main():
    torch.set_num_threads(os.cpu_count())
    ref_unet, ref_utils, ref_eval = ref_harness.import_reference()
    only = set(sys.argv[1:])                    # e.g. `python oracle/gen_golden.py amft` rewrites one family
    steps = [("memory", lambda: gen_memory(ref_unet)), ("amft", lambda: gen_amft(ref_unet)),
             ("psnr", lambda: gen_psnr(ref_utils)), ("scores", lambda: gen_scores(ref_eval)), ("records", gen_records),
             ("generator", lambda: gen_generator(ref_unet)), ("preprocess", gen_preprocess), ("losses", gen_losses), ("objectives", gen_objectives)]
    for name, fn in steps:
        if not only or name in only:
            fn()
    print("fixtures written to", GOLD)


if __name__ == "__main__":
    main()
