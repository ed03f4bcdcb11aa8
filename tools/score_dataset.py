"""Sharded inference scoring of a whole (synthetic) dataset: sub-videos LPT-partitioned over ranks, per-frame records
gathered on rank 0 (reference pickle schema), regularity scores + AUC computed there.  Exercises rows a12/a13/(e).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/score_dataset.py --dataset ped2 --size 128
"""
import argparse, json, os, pickle, sys, tempfile, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import dist as adist, synth

LENGTHS = {"ped2": synth.PED2_VIDEO_LENGTHS,
           "avenue": [1439, 1211, 923, 947, 1007, 1283, 605, 36, 1175, 841, 472, 1271, 549, 507, 1001, 740, 426, 294, 248, 273, 76]}


def video(v, T, S, dev):
    g = torch.Generator().manual_seed(1000 + v)
    base = torch.rand((1, 3, S, S), generator=g) * 2 - 1
    rgb = (base + 0.05 * torch.randn((T, 3, S, S), generator=g)).clamp(-1, 1)
    op = 0.02 * torch.randn((T - 1, 2, S, S), generator=g)
    return rgb.to(dev), op.to(dev)


NATIVE = {"ped2": (240, 360), "avenue": (360, 640), "shanghaitech": (480, 856)}


def native_video(v, T, hw):
    """Decoded frames (uint8 BGR) and .flo payloads of one sub-video in pinned host memory -- what a decoder thread hands
    over; everything after the decode (resize, colour order, normalisation) happens on the GPU."""
    rng = np.random.default_rng(1000 + v)
    # a new scene every 8 frames with its own contrast, so that PSNR and commit records vary along the video (a constant
    # record makes the reference's per-video min-max normalisation 0/0)
    scenes = rng.integers(0, 256, ((T + 7) // 8,) + hw + (3,), dtype=np.int16)
    gain = rng.uniform(0.2, 1.0, ((T + 7) // 8, 1, 1, 1))
    base = (128 + (scenes - 128) * gain).astype(np.int16)[np.arange(T) // 8]
    frames = np.clip(base + rng.integers(-12, 13, (T,) + hw + (3,), dtype=np.int16), 0, 255).astype(np.uint8)
    flows = (rng.standard_normal((T - 1,) + hw + (2,)) * 2).astype(np.float32)
    return torch.from_numpy(frames).pin_memory(), torch.from_numpy(flows).pin_memory()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--native", action="store_true", help="host uint8 frames at the dataset's native size: H2D + GPU "
                    "preprocessing inside the timed region, frames resized to 256x256")
    ap.add_argument("--dataset", default="ped2")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(20200525)
    g = A.get_twostream().to(dev).eval()
    lengths = LENGTHS[args.dataset]
    mine = adist.lpt_partition(lengths, world)[rank]
    scorer = A.VideoScorer(g, batch=args.batch, graph=args.graph)
    if args.graph:                                   # capture outside the timed region
        rgb0, op0 = video(0, args.batch + 4, 256 if args.native else args.size, dev)
        scorer.score_video(rgb0, op0)
    host = {v: native_video(v, lengths[v], NATIVE[args.dataset]) for v in mine} if args.native else None
    torch.cuda.synchronize()
    t0 = time.time()
    local_rec = {}
    h2d = 0
    for v in mine:
        if args.native:
            fr, fl = host[v]
            h2d += fr.numel() + fl.numel() * 4
            rgb = A.preprocess_frames(fr.to(dev, non_blocking=True), (256, 256))
            op = A.preprocess_flow(fl.to(dev, non_blocking=True), (256, 256))
        else:
            rgb, op = video(v, lengths[v], args.size, dev)
        local_rec[v] = scorer.score_video(rgb, op)
    torch.cuda.synchronize()
    dt = torch.tensor([time.time() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    rec = adist.gather_records(local_rec, len(lengths))
    if rank == 0:
        rec["dataset"] = args.dataset
        rng = np.random.RandomState(3)
        labels = [(rng.rand(n) < 0.3).astype(np.int8) for n in lengths]
        with tempfile.TemporaryDirectory() as td:
            pk = os.path.join(td, args.dataset)
            pickle.dump(rec, open(pk, "wb"))
            res = A.evaluate("img_pred_fea_comm_rgb_auc", pk, A.LAM_MAP.get(args.dataset, (0.05, 0.6)), gt_labels=labels)
        scored = sum(n - 4 for n in lengths)
        print(json.dumps({"dataset": args.dataset, "n_gpus": world, "videos": len(lengths), "frames_scored": scored,
                          "seconds": float(dt), "frames_per_s_end_to_end_with_unet": scored / float(dt), "auc_synthetic_labels": res["auc"],
                          "input": ("host uint8 %dx%d frames + fp32 flow, H2D and GPU preprocessing timed" % NATIVE[args.dataset])
                          if args.native else ("device-resident fp32 %dx%d frames" % (args.size, args.size)),
                          "h2d_bytes_rank0": h2d, "generator": getattr(g, "engine", "cudnn"),
                          "loads": [sum(lengths[i] for i in p) for p in adist.lpt_partition(lengths, world)]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
