"""One forward + backward of the fused generator objective (Twostream_vq_Loss) at 64 frames of 256 x 256, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/objective_once.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = {k: v.cuda() for k, v in synth.objective_inputs(dict(seed=5, b=n, h=256, w=256, hd=34, wd=34)).items()}
for k in ("rgb_out", "op_out", "d_gen"):
    t[k].requires_grad_(True)
fn = A.Twostream_vq_Loss(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
for _ in range(2):
    fn(t["flow_pred"], t["flow_gt"], t["rgb_out"], t["rgb_tgt"], t["op_out"], t["op_tgt"], t["latent"], t["d_gen"]).backward()
torch.cuda.synchronize()
print("g_loss", fn.g_loss)
