"""Fused frame losses vs the same losses written with torch ops (what the reference runs), 256x256 RGB frames."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
dev = "cuda:0"
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
for n in (8, 64, 256):
    gen, gt = (t.to(dev) for t in synth.frames(3, n))
    gen.requires_grad_(True)

    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    def ours():
        gen.grad = None
        li, lg = A.frame_losses(gen, gt)
        (li + lg).backward()

    def torch_ops():
        gen.grad = None
        (O.intensity_loss(gen, gt) + O.gradient_loss(gen, gt)).backward()

    a, b = timed(ours), timed(torch_ops)
    by = gen.numel() * 4 * 5            # fwd reads gen+gt, bwd reads gen+gt and writes grad
    print(json.dumps({"frames": n, "fused_fwd_bwd_ms": a, "torch_ops_fwd_bwd_ms": b, "speedup": b / a,
                      "fused_GBps": by / a / 1e6, "frac_of_hbm_peak": by / a / 1e6 / hbm}), flush=True)

# element-wise training objectives (Flow_Loss on [n,2,256,256] flows; Discriminate_Loss / Adversarial_Loss on [n,1,34,34] maps
# of the reference discriminator) and the whole generator objective Twostream_vq_Loss against the same sums in torch ops
LAM = dict(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
for n in (8, 64):
    t = {k: v.to(dev) for k, v in synth.objective_inputs(dict(seed=5, b=n, h=256, w=256, hd=34, wd=34)).items()}
    for k in ("flow_pred", "rgb_out", "op_out", "d_gen", "d_real"):
        t[k].requires_grad_(True)

    def clear():
        for v in t.values():
            v.grad = None

    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    fl, dl, gl = A.Flow_Loss(), A.Discriminate_Loss(), A.Twostream_vq_Loss(**LAM)

    def flow_ours(): clear(); fl(t["flow_pred"], t["flow_gt"]).backward()
    def flow_torch(): clear(); O.flow_loss(t["flow_pred"], t["flow_gt"]).backward()
    def dis_ours(): clear(); dl(t["d_real"], t["d_gen"]).backward()
    def dis_torch(): clear(); O.discriminate_loss(t["d_real"], t["d_gen"]).backward()
    def obj_ours():
        clear(); gl(t["flow_pred"], t["flow_gt"], t["rgb_out"], t["rgb_tgt"], t["op_out"], t["op_tgt"], t["latent"], t["d_gen"]).backward()
    def obj_torch():
        clear()
        loss, parts = O.twostream_vq_loss(LAM, t["flow_pred"], t["flow_gt"], t["rgb_out"], t["rgb_tgt"], t["op_out"], t["op_tgt"],
                                          t["latent"], t["d_gen"])
        loss.backward()
        [v.item() for v in parts.values()]                      # the reference's per-scalar .item() reads (loss_zoo.py:341-348)

    fa, fb, da, db, oa, ob = (timed(f) for f in (flow_ours, flow_torch, dis_ours, dis_torch, obj_ours, obj_torch))
    fby = t["flow_pred"].numel() * 4 * 5
    print(json.dumps({"frames": n, "flow_loss_fused_fwd_bwd_ms": fa, "flow_loss_torch_ops_ms": fb, "flow_speedup": fb / fa,
                      "flow_fused_GBps": fby / fa / 1e6, "flow_frac_of_hbm_peak": fby / fa / 1e6 / hbm,
                      "discriminate_loss_fused_ms": da, "discriminate_loss_torch_ops_ms": db, "discriminate_speedup": db / da,
                      "twostream_vq_loss_fused_fwd_bwd_ms": oa, "twostream_vq_loss_torch_ops_ms": ob, "objective_speedup": ob / oa}),
          flush=True)
