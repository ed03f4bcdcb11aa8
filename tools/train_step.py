"""BASELINE configs[3]: joint training step of the two-stream generator, data-parallel (one process per GPU, NCCL).

The hot-path modules run their training kernels (memory: fused forward + closed-form backward + EMA bank update with
globally all-reduced assignment statistics; AMFT: batch-stat BN, tcgen05 dgrad/wgrad); the U-Net convolutions are stock
cuDNN layers (unchanged host code).  Losses follow the reference step structure (Code/run_helper/train_helper.py:291-343,
Code/models/losses/loss_zoo.py:307-350) without the parts that need unavailable weights (FlowNet2-SD, discriminator):
intensity L2 on the rgb prediction, L1 on the flow prediction, lam_latent * (rgb_diff + op_diff).

`--gan` runs the reference's adversarial step structure (train_helper.py:318-339): the discriminator
(`PixelDiscriminator(3, [128, 256, 512, 512])`, cuDNN layers) is updated on (target, detached prediction) with the fused
`Discriminate_Loss`, then the generator on `Twostream_vq_Loss` (fused adversarial / flow / intensity / gradient objectives,
one host read of its eight scalars).  The discriminator update comes first here: the reference back-propagates the
generator objective through discriminator weights its optimizer has already stepped, which current autograd refuses.
With `--flownet` the frozen flow estimator (`FlowNet2SD` host mirror on cuDNN layers, random initial weights: no checkpoint
exists offline) produces the two flows `Flow_Loss` compares, exactly as train_helper.py:309-322 does (detached, inputs
rescaled to (0, 255), outputs / 255); without it `Flow_Loss` is fed the flow-stream prediction / target -- tensors of the
same shape and role.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py --steps 10 --batch 8
"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import dist as adist, functions as F_


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)       # per GPU (reference script: batch 8, training_com.sh:21)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--sync-bn", action="store_true", help="global-batch BatchNorm statistics (single-GPU semantics)")
    ap.add_argument("--gan", action="store_true", help="adversarial step structure: discriminator + Twostream_vq_Loss")
    ap.add_argument("--flownet", action="store_true", help="with --gan: frozen FlowNet2SD (random weights) feeds Flow_Loss")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        adist.install_stats_allreduce()
    torch.manual_seed(20200525)                           # same init on every rank (reference seeds at import, unet.py:4)
    g = A.get_twostream().to(dev).train()
    if args.sync_bn:
        g = adist.install_sync_bn(g)
    opt = torch.optim.Adam(g.parameters(), lr=2e-4)
    params = [p for p in g.parameters()]
    gen = torch.Generator().manual_seed(77 + rank)
    B, S = args.batch, args.size
    rgb = (torch.rand((B, 5, 3, S, S), generator=gen) * 2 - 1).to(dev)
    op = (torch.randn((B, 4, 2, S, S), generator=gen) * 0.02).to(dev)
    rgb_in, rgb_tgt = rgb[:, :-1].flatten(1, 2), rgb[:, -1]
    op_in, op_tgt = op[:, :-1].flatten(1, 2), op[:, -1]
    lam_latent = 1.0
    if args.gan:
        D = A.PixelDiscriminator(3, [128, 256, 512, 512], use_norm=False).to(dev).train()
        opt_d = torch.optim.Adam(D.parameters(), lr=2e-5)
        d_params = [p for p in D.parameters()]
        g_loss_fn = A.Twostream_vq_Loss(lam_adv=0.05, lam_gdl=1.0, lam_flow=2.0, lam_lp=1.0, lam_latent=0.1, lam_lp_op=2.0)
        d_loss_fn = A.Discriminate_Loss()
        flow_net = A.FlowNet2SD().to(dev).eval() if args.flownet else None
        rgb_last = rgb[:, -1]                                   # `rgb_input_last` of train_helper.py:299

    def flows_of(pr):
        if flow_net is None:
            return None
        with torch.no_grad():
            pair = lambda second: (torch.stack([rgb_last, second], 2) * 0.5 + 0.5) * 255.0
            return flow_net(pair(pr.detach())) / 255.0, flow_net(pair(rgb_tgt)) / 255.0

    def gan_step():
        pr, po, diffs, _ = g(rgb_in, op_in)
        opt_d.zero_grad(set_to_none=True)                       # (1) discriminator on (real, detached fake)
        d_loss = d_loss_fn(D(rgb_tgt), D(pr.detach()))
        d_loss.backward()
        adist.allreduce_gradients(d_params)
        opt_d.step()
        opt.zero_grad(set_to_none=True)                         # (2) generator; the discriminator only carries the gradient
        for p in d_params:
            p.requires_grad_(False)
        fl = flows_of(pr)
        flow_pred, flow_gt = fl if fl is not None else (po, op_tgt)
        loss = g_loss_fn(flow_pred, flow_gt, pr, rgb_tgt, po, op_tgt, diffs, D(pr))
        loss.backward()
        for p in d_params:
            p.requires_grad_(True)
        adist.allreduce_gradients(params)
        opt.step()
        return loss.detach().reshape(())

    def step():
        if args.gan:
            return gan_step()
        opt.zero_grad(set_to_none=True)
        pr, po, (d_rgb, d_op), _ = g(rgb_in, op_in)
        # the reference's generator objective without the adversarial / FlowNet terms (loss_zoo.py:124-126, 190-192): intensity +
        # gradient loss on the frame, intensity loss on the flow, commit terms -- the image-space losses from the fused kernels
        l_int, l_gd = A.frame_losses(pr, rgb_tgt)
        l_int_op, _ = A.frame_losses(po, op_tgt)
        loss = l_int + l_gd + 2.0 * l_int_op + lam_latent * (d_rgb + d_op).sum()
        loss.backward()
        adist.allreduce_gradients(params)
        opt.step()
        return loss.detach()

    losses = []
    for _ in range(args.warmup):
        losses.append(step())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses.append(step())
    e1.record()
    torch.cuda.synchronize()
    F_.check_pipeline_watchdog()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # consistency: banks and parameters must be identical on every rank after the step
    bank = g.rgb.vq_down3.quan.quantize.embed.detach().clone()
    w = g.bridge.O2F.conv[0].weight.detach().clone()
    bank_max_dev = torch.zeros(1, device=dev, dtype=torch.float64)
    w_max_dev = torch.zeros(1, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref_b, ref_w = bank.clone(), w.clone()
        dist.broadcast(ref_b, 0); dist.broadcast(ref_w, 0)
        bank_max_dev = (bank - ref_b).abs().max().double().reshape(1)
        w_max_dev = (w - ref_w).abs().max().double().reshape(1)
        dist.all_reduce(bank_max_dev, op=dist.ReduceOp.MAX); dist.all_reduce(w_max_dev, op=dist.ReduceOp.MAX)
    if rank == 0:
        ls = [float(l) for l in losses]
        print(json.dumps({"what": "joint training step (generator fwd+bwd, EMA stats + gradient all-reduce, Adam)" + (" + discriminator update, Twostream_vq_Loss" + (", frozen FlowNet2SD x2" if args.flownet else "") if args.gan else ""),
                          "n_gpus": world, "batch_per_gpu": B, "batch_norm": "global batch (all-reduced sums)" if args.sync_bn else "per rank", "steps": args.steps, "ms_per_step": float(ms) / args.steps,
                          "frames_per_s": world * B * args.steps / (float(ms) * 1e-3), "loss_first": ls[0], "loss_last": ls[-1],
                          "finite": all(l == l and abs(l) < 1e30 for l in ls),
                          "bank_max_abs_dev_across_ranks": float(bank_max_dev), "weight_max_abs_dev_across_ranks": float(w_max_dev)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
