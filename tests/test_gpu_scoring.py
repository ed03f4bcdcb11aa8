"""GPU parity: batched PSNR, bit-exact score reduction, record assembly through the scorer."""
import numpy as np
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, scoring
from conftest import load_golden, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_psnr_vs_golden():
    c, g = load_golden("psnr")
    gen, gt = synth.frames(c["seed"], c["b"], c["c"], c["h"], c["w"])
    per = A.psnr_per_frame(gen.to(DEV), gt.to(DEV))
    assert_close(per.cpu(), g["per_frame"], 1e-5, "psnr.per_frame")
    assert_close(A.psnr_error(gen.to(DEV), gt.to(DEV)).cpu(), g["batch_mean"], 1e-5, "psnr.batch")
    one = A.psnr_error(gen[:1].to(DEV), gt[:1].to(DEV))
    assert one.dim() == 0
    assert_close(one.cpu().reshape(1), g["per_frame"][:1], 1e-5, "psnr.single")


@pytest.mark.parametrize("shape", [(64, 3, 256, 256), (3, 3, 37, 41), (1, 1, 1, 1), (5, 2, 256, 256)])
def test_psnr_shapes_vs_oracle(shape):
    gen, gt = synth.frames(41, *shape)
    per = A.psnr_per_frame(gen.to(DEV), gt.to(DEV))
    ref = O.psnr_per_frame(gen, gt)
    assert_close(per.cpu(), ref, 1e-5, f"psnr{shape}")


def test_psnr_unaligned_view():
    gen, gt = synth.frames(42, 4, 3, 33, 35)
    g1, t1 = gen.to(DEV)[:, :, 1:, 2:], gt.to(DEV)[:, :, 1:, 2:]
    assert_close(A.psnr_per_frame(g1, t1).cpu(), O.psnr_per_frame(gen[:, :, 1:, 2:], gt[:, :, 1:, 2:]), 1e-5, "psnr.view")


@pytest.mark.parametrize("ds", ["ped2", "avenue", "shanghaitech"])
def test_score_reduce_bit_exact(ds):
    c, g = load_golden("scores_" + ds)
    offs = np.concatenate([[0], np.cumsum(g["lengths"])])
    img = [g["img"][offs[i]:offs[i + 1]] for i in range(len(g["lengths"]))]
    fea = [g["fea"][offs[i]:offs[i + 1]] for i in range(len(g["lengths"]))]
    s = A.score_reduce(img, fea, tuple(c["lam"]))
    assert s.dtype == np.float32 and s.shape == g["scores"].shape
    assert np.array_equal(s, g["scores"]), "max abs diff %g" % float(np.abs(s - g["scores"]).max())
    assert np.array_equal(img[0], g["img"][:offs[1]]), "caller records must not be normalised in place"


def test_score_reduce_ragged_and_constant():
    rng = np.random.RandomState(3)
    lens = [5, 6, 40, 1439, 9]
    img = [rng.rand(n).astype(np.float32) * 30 + 10 for n in lens]
    fea = [rng.rand(n).astype(np.float32) * 1e-5 for n in lens]
    for lam in [(0.0, 0.0), (0.5, 1.0), (0.13, 0.6)]:
        assert np.array_equal(A.score_reduce(img, fea, lam), O.score_reduce(img, fea, lam))


def test_evaluate_matches_sklearn_on_reference_records(tmp_path):
    import pickle
    c, g = load_golden("scores_ped2")
    offs = np.concatenate([[0], np.cumsum(g["lengths"])])
    nv = len(g["lengths"])
    rec = {"dataset": "ped2",
           "rgb_img_pred_records": [g["img"][offs[i]:offs[i + 1]].copy() for i in range(nv)],
           "rgb_fea_comm_records": [g["fea"][offs[i]:offs[i + 1]].copy() for i in range(nv)],
           "op_img_pred_records": [], "op_fea_comm_records": []}
    pk = tmp_path / "ped2"
    pickle.dump(rec, open(pk, "wb"))
    labels = [g["labels"][offs[i]:offs[i + 1]] for i in range(nv)]
    ret = A.evaluate("img_pred_fea_comm_rgb_auc", str(pk), tuple(c["lam"]), gt_labels=labels)
    assert ret["auc"] == round(float(g["auc"]), 3) and ret["optimal_loss"] == str(pk)


def test_video_scorer_reproduces_reference_loop_semantics():
    """Row a12: the batched, sync-free scorer must write the records the reference loop writes
    (Code/run_helper/test_helper.py:408-475): per-frame PSNR, the commit scalar of each 16-clip group, back-filled head,
    op-stream tail copy -- independent of the batch size used on the device."""
    torch.manual_seed(0)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = A.get_twostream().to(DEV).eval()
        T, H = 41, 64                                        # 37 clips -> groups of 16, 16, 5
        gen = torch.Generator().manual_seed(5)
        rgb = (torch.rand((T, 3, H, H), generator=gen) * 2 - 1).to(DEV)
        op = (torch.randn((T - 1, 2, H, H), generator=gen) * 0.02).to(DEV)
        rec = scoring.VideoScorer(g, batch=64).score_video(rgb, op)
        rec7 = scoring.VideoScorer(g, batch=7).score_video(rgb, op)
        rec_graph = scoring.VideoScorer(g, batch=16, graph=True).score_video(rgb, op)    # 2 replays + 1 eager batch
        for kk in ("rgb_img_pred", "rgb_fea_comm", "op_fea_comm"):
            np.testing.assert_allclose(rec_graph[kk], scoring.VideoScorer(g, batch=16).score_video(rgb, op)[kk], rtol=1e-5)
        # the reference's own procedure: batches of 16 clips, psnr_error per frame, the batch-level diff per frame
        n_clips = T - 4
        psnr_ref = np.empty(n_clips, np.float32)
        commit_ref = np.empty(n_clips, np.float32)
        commit_op_ref = np.empty(n_clips, np.float32)
        with torch.no_grad():
            for c0 in range(0, n_clips, 16):
                idx = torch.arange(c0, min(n_clips, c0 + 16), device=DEV)
                rgb_in = torch.stack([rgb[idx + t] for t in range(4)], 1).flatten(1, 2)
                op_in = torch.stack([op[idx + t] for t in range(3)], 1).flatten(1, 2)
                pred, _, (d_rgb, d_op), _ = g(rgb_in, op_in)
                for i in range(idx.numel()):
                    psnr_ref[c0 + i] = float(A.psnr_error(pred[i:i + 1], rgb[idx[i] + 4][None]))
                    commit_ref[c0 + i] = float(d_rgb)
                    commit_op_ref[c0 + i] = float(d_op)
        img = np.concatenate([np.full(4, psnr_ref[0], np.float32), psnr_ref])
        fea = np.concatenate([np.full(4, commit_ref[0], np.float32), commit_ref])
        assert rec["rgb_img_pred"].shape == (T,) and rec["op_fea_comm"].shape == (T,)
        for r in (rec, rec7):
            assert_close(r["rgb_img_pred"], img, 1e-3, "records.psnr")
            assert_close(r["rgb_fea_comm"], fea, 1e-3, "records.commit")
            op_fea = np.concatenate([np.full(3, commit_op_ref[0], np.float32), commit_op_ref, commit_op_ref[-1:]])
            assert_close(r["op_fea_comm"], op_fea, 1e-3, "records.op_commit")
        assert len(set(np.round(rec["rgb_fea_comm"][4:], 12))) == 3       # one commit value per 16-clip group
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_graphed_path_matches_eager():
    """CUDA-graph replay of memory modules + AMFT + PSNR must reproduce the eager results bit for bit."""
    from ammcnet_aaai2021_b200 import functions as F_
    C, D, M, k, b = 512, 64, 256, 2, 4
    p = synth.path_params(6, C, D, M, k)
    mods = {}
    for s in ("rgb", "op"):
        m = A.enc_quan_dec_res_topk(C, D, M, k=k)
        pre = s + ".vq_down3."
        m.load_state_dict({kk[len(pre):]: v for kk, v in p.items() if kk.startswith(pre)}, strict=True)
        mods[s] = m.to(DEV).eval()
    br = A.bridge(in_c=C)
    br.load_state_dict({kk[len("bridge."):]: v for kk, v in p.items() if kk.startswith("bridge.")}, strict=True)
    br = br.to(DEV).eval()

    def step(xr, xo, gen, gt):
        o_r, _, _ = mods["rgb"](xr)
        o_o, _, _ = mods["op"](xo)
        yr, yo = br(o_r, o_o)
        return yr, yo, F_.psnr_per_frame(gen, gt), mods["rgb"].quan.quantize.last_sse_frame

    ins = [synth.features(51, b, C, 32, 32).to(DEV), synth.features(52, b, C, 32, 32).to(DEV)]
    ins += [t.to(DEV) for t in synth.frames(53, b, 3, 64, 64)]
    with torch.no_grad():
        eager = [t.clone() for t in step(*ins)]
    g = A.GraphedPath(step, ins)
    outs = g(*ins)
    torch.cuda.synchronize()
    for a, e in zip(outs, eager):
        assert torch.equal(a, e)
    ins2 = [synth.features(61, b, C, 32, 32).to(DEV), synth.features(62, b, C, 32, 32).to(DEV)]
    ins2 += [t.to(DEV) for t in synth.frames(63, b, 3, 64, 64)]
    with torch.no_grad():
        eager2 = [t.clone() for t in step(*ins2)]
    outs2 = g(*ins2)
    torch.cuda.synchronize()
    for a, e in zip(outs2, eager2):
        assert torch.equal(a, e)
    F_.check_pipeline_watchdog()


@pytest.mark.parametrize("ds", ["ped2", "avenue", "shanghaitech"])
def test_gpu_roc_auc_matches_sklearn_on_reference_records(ds):
    """Next-row: ROC-AUC on the device equals sklearn's on the reference's recorded score records."""
    from ammcnet_aaai2021_b200 import functions as F_
    c, g = load_golden("scores_" + ds)
    offs = np.concatenate([[0], np.cumsum(g["lengths"])])
    nv = len(g["lengths"])
    labels = np.concatenate([g["labels"][offs[i] + 4:offs[i + 1]] for i in range(nv)]).astype(np.int8)
    scores = torch.from_numpy(g["scores"]).to(DEV)
    auc = float(F_.roc_auc_device(scores, torch.from_numpy(labels).to(DEV), pos_label=0).cpu())
    assert abs(auc - float(g["auc"])) < 1e-12, (auc, float(g["auc"]))
    img = [g["img"][offs[i]:offs[i + 1]] for i in range(nv)]
    fea = [g["fea"][offs[i]:offs[i + 1]] for i in range(nv)]
    lab = [g["labels"][offs[i]:offs[i + 1]] for i in range(nv)]
    s, a = scoring.score_and_auc_device(img, fea, tuple(c["lam"]), lab)
    assert np.array_equal(s, g["scores"]) and abs(a - float(g["auc"])) < 1e-12


@pytest.mark.parametrize("T", [1, 2, 7, 2048, 2049, 5000, 70000])
def test_gpu_roc_auc_ties_and_sizes(T):
    from ammcnet_aaai2021_b200 import functions as F_
    rng = np.random.RandomState(T)
    scores = np.round(rng.randn(T), 1).astype(np.float32)           # heavy ties, negatives, zeros
    labels = (rng.rand(T) < 0.4).astype(np.int8)
    got = float(F_.roc_auc_device(torch.from_numpy(scores).to(DEV), torch.from_numpy(labels).to(DEV), 0).cpu())
    ref = O.roc_auc(labels, scores, pos_label=0)
    if np.isnan(ref):
        assert np.isnan(got)
    else:
        assert abs(got - ref) < 1e-12, (got, ref)
