#!/usr/bin/env python
"""Benchmark of the AMMC-Net memory + AMFT + score hot path (BASELINE.json metric, config #2 shape at N=1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the path over a batch of 64 synthetic frames per GPU: both memory modules on
[64,512,32,32] bottleneck features (D=64, M=256, k=2), the AMFT block, and the rgb PSNR of 64 3x256x256 frame
pairs.  Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same
metric through the public module API with pinned HOST buffers copied in (and scores copied out) every step.
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU ops, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (memory+AMFT+score path)"
UNIT = "frames/s"
C, D, M, K_TOP, HW = 512, 64, 256, 2, 32     # M is overridden by --items
FRAME = (3, 256, 256)


def workload_desc(batch, precision):
    return {
        "workload": "BASELINE configs[1]: AMMC-Net ped2-shape inference path on synthetic frames, batch %d per GPU: "
                    "2 memory modules on [%d,512,32,32] (D=64, M=%d, k=2) + AMFT bridge(512) + rgb PSNR on "
                    "[%d,3,256,256]" % (batch, batch, M, batch),
        "batch_per_gpu": batch,
        "arithmetic": ("split-bf16 x3 tensor-core passes, fp32 accumulate (fp32-parity mode)" if precision == 3
                       else "single bf16 tensor-core pass, fp32 accumulate") + "; memory addressing, PSNR in fp32",
        "l2_policy": "inputs+intermediates per step (~0.9 GB) exceed the 126 MB L2; no explicit flush",
        "launch": "one CUDA-graph replay per step (eager with --no-graph)",
        "sharding": "clips data-parallel, replicated bank and weights; per-frame scores all-gathered each step",
    }


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def _oracle_step_fn(frames):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ammc_oracle as O          # the ONLY place bench.py touches oracle/: the timed CPU baseline
    from ammcnet_aaai2021_b200 import synth
    p = synth.path_params(1, C, D, M, K_TOP)
    xr, xo = synth.features(11, frames, C, HW, HW), synth.features(12, frames, C, HW, HW)
    gen, gt = synth.frames(13, frames, *FRAME)

    def step():
        with torch.no_grad():
            return O.path_forward(xr, xo, gen, gt, p, K_TOP)
    return step


def cpu_baseline(budget_s=12.0, frames_per_call=4, max_frames=64):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _oracle_step_fn(frames_per_call)
    step()                                           # warm-up (thread pools, mkldnn primitives)
    done, t0 = 0, time.perf_counter()
    while done < max_frames and (time.perf_counter() - t0) < budget_s:
        step()
        done += frames_per_call
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d frames of the same workload (oracle/ammc_oracle.py path_forward, torch CPU fp32, %d threads, "
                      "%d frames per call), %.1f s" % (done, cores, frames_per_call, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = 4
    step = _oracle_step_fn(frames)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = frames * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_desc(64, 3), sample="each step = %d frames of the workload on the host CPU" % frames),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps x %d frames, oracle port of the reference's torch-CPU op sequence "
                                   "(the Python reference cannot travel to the GPU box)" % (args.steps, frames)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.proc = subprocess.Popen(["nvidia-smi", "-i", sel, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    @staticmethod
    def _ts(s):
        import datetime
        try:
            return datetime.datetime.strptime(s.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.file.read().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            ts = self._ts(f[0])
            # keep the samples taken while the timed loop was running (nvidia-smi stamps are host local time)
            if ts is not None and self.t_begin is not None and not (self.t_begin - 0.02 <= ts <= self.t_end + 0.02):
                continue
            f = f[1:]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.file.name)
        if sm:
            sm_sorted = sorted(sm)
            out.update(sm_mhz=sm_sorted[len(sm_sorted) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this same command (profiles/, summarised by tools/ncu_summary.py); None when the file is absent."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01_ncu_conv_igemm_pair_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0) for r in rows[2:] if r]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def run_ours(args):
    import torch.distributed as dist
    import ammcnet_aaai2021_b200 as A
    from ammcnet_aaai2021_b200 import functions as F_, synth
    from ammcnet_aaai2021_b200 import dist as adist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    prec = args.precision
    if args.no_pair:
        F_.set_conv_pair_mode(False)
    if args.pair_unfused:
        F_.set_conv_pair_mode(3)

    # ---- modules with random-init weights of the shipped architecture (replicated on every rank) ----------
    p = synth.path_params(1, C, D, M, K_TOP)
    mem = {}
    for s in ("rgb", "op"):
        m = A.enc_quan_dec_res_topk(C, D, M, k=K_TOP)
        pre = s + ".vq_down3."
        m.load_state_dict({k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}, strict=True)
        mem[s] = m.to(dev).eval()
    amft = A.bridge(in_c=C, precision=prec)
    amft.load_state_dict({k[len("bridge."):]: v for k, v in p.items() if k.startswith("bridge.")}, strict=True)
    amft = amft.to(dev).eval()

    def local_step(xr, xo, gen, gt):
        with torch.no_grad():
            if args.no_streams:
                (o_r, d_r, _), (o_o, d_o, _), ps = mem["rgb"](xr), mem["op"](xo), F_.psnr_per_frame(gen, gt)
            else:       # the two memory modules and the PSNR are independent until the AMFT block: three streams
                (o_r, d_r, _), (o_o, d_o, _), ps = F_.concurrently(lambda: mem["rgb"](xr), lambda: mem["op"](xo),
                                                                   lambda: F_.psnr_per_frame(gen, gt))
            yr, yo = amft(o_r, o_o)
            commit = mem["rgb"].quan.quantize.last_sse_frame
            scores = torch.stack([ps, commit])                    # per-frame (psnr, commit partial)
        return yr, yo, scores

    def exchange(out):
        yr, yo, scores = out
        if world > 1:
            scores = adist.all_gather_scores(scores)               # inference exchange step: scores only (NCCL)
        return yr, yo, scores

    def path_step(xr, xo, gen, gt):                                # eager form (also used for the per-kernel timing)
        return exchange(local_step(xr, xo, gen, gt))

    # ---- synthetic inputs: host (pinned) and device copies -------------------------------------------------
    xr_h = synth.features(1234 + rank, B, C, HW, HW).pin_memory()
    xo_h = synth.features(4321 + rank, B, C, HW, HW).pin_memory()
    gen_h, gt_h = synth.frames(99 + rank, B, *FRAME)
    gen_h, gt_h = gen_h.pin_memory(), gt_h.pin_memory()
    xr, xo, gen, gt = (t.to(dev, non_blocking=True) for t in (xr_h, xo_h, gen_h, gt_h))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None      # started before warm-up so it is sampling by the time we time
    use_graph = not args.no_graph
    graphed = None
    if use_graph:
        from ammcnet_aaai2021_b200.graphs import GraphedPath
        graphed = GraphedPath(local_step, [xr, xo, gen, gt])
        step_fn = lambda: exchange(graphed.replay())               # inputs already sit in the captured buffers
    else:
        step_fn = lambda: path_step(xr, xo, gen, gt)
    for _ in range(max(args.warmup, 3)):
        step_fn()
    barrier()

    # ---- timed region (device-resident inputs) --------------------------------------------------------------
    F_.LAUNCHES["count"] = 0
    path_step(xr, xo, gen, gt)
    launches_per_step = F_.LAUNCHES["count"]                     # kernels of ours in one step (a graph replays the same)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step_fn()
    e1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    # dominant-kernel timing: CUDA events around every conv launch of the same K steps issued eagerly (events cannot be
    # recorded inside a replayed graph), same stream, right after the timed region
    F_.PROFILE["on"] = True
    F_.PROFILE["events"].clear()
    for _ in range(args.steps):
        path_step(xr, xo, gen, gt)
    torch.cuda.synchronize()
    F_.PROFILE["on"] = False
    conv_ms = [s.elapsed_time(e) for (s, e) in F_.PROFILE["events"] if s.elapsed_time(e) > 0.2]   # 3x3 convs only
    F_.PROFILE["events"].clear()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- component breakdown (rank 0, informational) -------------------------------------------------------
    def timed(fn, n=5):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b_.record(); torch.cuda.synchronize()
        return a.elapsed_time(b_) / n

    breakdown = None
    if rank == 0:
        with torch.no_grad():
            o_r, _, _ = mem["rgb"](xr)
            o_o, _, _ = mem["op"](xo)
            breakdown = {
                "memory_module_x2_ms": timed(lambda: (mem["rgb"](xr), mem["op"](xo))),
                "amft_ms": timed(lambda: amft(o_r, o_o)),
                "psnr_ms": timed(lambda: F_.psnr_per_frame(gen, gt)),
            }
            del o_r, o_o

    # ---- e2e: host buffers in, scores out, every step (double-buffered copies on a side stream) -------------
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [[torch.empty_like(t, device=dev) for t in (xr_h, xo_h, gen_h, gt_h)] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    h2d_bytes = sum(t.numel() * 4 for t in (xr_h, xo_h, gen_h, gt_h))
    scores_h = torch.empty((2, B) if world == 1 else (world, 2, B), dtype=torch.float32).pin_memory()

    e2e_graphs = None
    if use_graph:
        e2e_graphs = [GraphedPath(local_step, bufs[i]) for i in range(2)]
        for i in range(2):
            bufs[i] = e2e_graphs[i].static_inputs                 # copy straight into the captured buffers

    def e2e_loop(n):
        cur = torch.cuda.current_stream(dev)
        for i in range(n + 1):
            if i < n:                                             # stage inputs of step i
                sl = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[sl])
                    for d_t, h_t in zip(bufs[sl], (xr_h, xo_h, gen_h, gt_h)):
                        d_t.copy_(h_t, non_blocking=True)
                    ready[sl].record(copy_stream)
            if i > 0:                                             # compute step i-1
                sl = (i - 1) & 1
                cur.wait_event(ready[sl])
                out = e2e_graphs[sl].replay() if use_graph else local_step(*bufs[sl])
                _, _, sc = exchange(out)
                scores_h.copy_(sc, non_blocking=True)
                freed[sl].record(cur)
        return scores_h

    for ev in freed:
        ev.record(torch.cuda.current_stream(dev))
    e2e_loop(2)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    e2e_loop(args.steps)
    a1.record()
    barrier()
    t = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_val = world * B * args.steps / (e2e_ms * 1e-3)

    # ---- bf16 single-pass variant, stated separately ---------------------------------------------------------
    variant = None
    if prec == 3:
        amft.precision = 1
        if use_graph:
            g1 = GraphedPath(local_step, [xr, xo, gen, gt])
            vstep = lambda: exchange(g1.replay())
        else:
            vstep = lambda: path_step(xr, xo, gen, gt)
        for _ in range(3):
            vstep()
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(args.steps):
            vstep()
        v1.record()
        barrier()
        t = torch.tensor([v0.elapsed_time(v1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        variant = {"amft_single_bf16_pass_frames_per_s": world * B * args.steps / (float(t.item()) * 1e-3),
                   "note": "bf16 variant: AMFT error ~1e-2 relative, outside the fp32 1e-3 parity bar"}
        amft.precision = 3

    if rank == 0:
        peaks, peak_src = load_peaks()
        conv_flops = 2.0 * B * HW * HW * C * 9 * C                # algorithmic, counted once whatever the pass count
        roof = None
        if conv_ms:
            avg = sum(conv_ms) / len(conv_ms)
            ach = conv_flops / (avg * 1e-3) / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
            roof = {"kernel": "conv_igemm_pair_kernel<FUSED3> (AMFT 3x3 conv, tcgen05 cta_group::2)", "bound": "tensor", "achieved": ach,
                    "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu_traffic_bytes(),
                    "traffic_note": "mean DRAM bytes per launch over the two convs of one branch, ncu --set full "
                                    "(profiles/r01_ncu_conv_igemm_pair_full_summary.csv); algorithmic operand+result bytes "
                                    "per launch: 2*67 MB planes in + 134 MB out (+134 MB residual) + 9.4 MB weights",
                    "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "avg_launch_ms": avg, "launches_timed": len(conv_ms),
                    "timing": "CUDA events around each 3x3 conv launch of K eagerly issued steps run right after the "
                              "timed region (events cannot be recorded inside a graph replay)",
                    "algorithmic_flops_per_launch": conv_flops,
                    "tensor_passes": prec, "executed_frac_of_peak": prec * ach / peak}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3->f32acc" if prec == 3 else "bf16->f32acc", "data": "synthetic",
            "config": workload_desc(B, prec),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": scores_h.numel() * 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "breakdown": breakdown, "variants": variant,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        if world == 1 and not args.no_generator:
            line["generator"] = generator_leg(dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def generator_leg(dev, batch=16, steps=5):
    """SURVEY section 8(f) rank 1, reported beside the path's metric (never part of `value`): eval-mode forward of the
    whole twostream generator on 256x256 frames, tcgen05 conv engine vs the same module with its U-Net layers on cuDNN."""
    import ammcnet_aaai2021_b200 as A
    from ammcnet_aaai2021_b200 import synth
    m = A.get_twostream()
    m.load_state_dict(synth.generator_params(3))
    m = m.to(dev).eval()
    rgb, op = (t.to(dev) for t in synth.generator_inputs(9, batch, 256, 256))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return batch * steps / (e0.elapsed_time(e1) * 1e-3)

    with torch.no_grad():
        eng = A.GeneratorEngine(m)
        graph = A.GraphedPath(eng, [rgb, op])
        ours = timed(graph.replay)
        m.engine = "cudnn"
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        cudnn_tf32 = timed(lambda: m(rgb, op))
        torch.backends.cudnn.allow_tf32 = False
        cudnn_fp32 = timed(lambda: m(rgb, op))
        torch.backends.cudnn.allow_tf32 = tf32
    del graph, eng, m
    torch.cuda.empty_cache()
    return {"value": ours, "unit": "frames/s", "workload": "whole twostream generator (unet.py:981-1007), eval, 256x256 "
            "frames, batch %d, split-bf16 x3 (fp32 parity), CUDA-graph replay, inputs resident" % batch,
            "algorithmic_tflops": ours * 187.4e9 / 1e12,
            "same_module_unet_on_cudnn_tf32": cudnn_tf32, "same_module_unet_on_cudnn_fp32": cudnn_fp32}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", type=int, default=3, choices=[1, 3])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-generator", action="store_true", help="skip the whole-generator leg (SURVEY 8(f) rank 1)")
    ap.add_argument("--items", type=int, default=256, help="memory bank size M (BASELINE configs[2] sweeps 256..2000)")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--pair-unfused", action="store_true", help="A/B: CTA-pair kernel streaming the K loop three times")
    ap.add_argument("--no-streams", action="store_true", help="A/B: issue the two memory modules and the PSNR on one stream")
    ap.add_argument("--no-pair", action="store_true", help="A/B: single-CTA conv kernel instead of the CTA-pair one")
    args = ap.parse_args()
    global M
    M = args.items
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
