"""Data-parallel plumbing for the path: one process per GPU, torch.distributed (NCCL on
the GPUs, gloo in CPU tests).  The path shards by sub-video / clip with a replicated bank and replicated weights,
so inference needs NO data-path collective -- only the per-frame score records are gathered at the end -- and
training needs (i) the gradient all-reduce and (ii) an all-reduce of the EMA assignment statistics BEFORE the
bank update, so every rank applies the identical update (stock DDP buffer broadcast would keep rank 0's only).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import modules as _modules


# --------------------------------------------------------------------------------------------------
# inference: shard sub-videos, gather records
# --------------------------------------------------------------------------------------------------
def lpt_partition(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time assignment of sub-videos to ranks (video lengths vary 36..1439 frames on avenue).
    Returns, per rank, the sorted list of video indices it owns.  Whole videos stay on one rank so the reference's
    16-clip commit groups (test_helper.py:414) are never split."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    owner: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        owner[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(o) for o in owner]


def gather_records(local: Dict[int, Dict[str, np.ndarray]], n_videos: int, group=None, dst: int = 0):
    """Each rank holds {video_index: {'rgb_img_pred': arr, 'rgb_fea_comm': arr, ...}}; rank `dst` receives the four
    record lists in video order (the reference pickle layout, test_helper.py:479-483); other ranks get None."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        gathered = [local]
    else:
        gathered = [None] * world if rank == dst else None
        dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    merged: Dict[int, Dict[str, np.ndarray]] = {}
    for part in gathered:
        merged.update(part)
    if sorted(merged) != list(range(n_videos)):
        raise RuntimeError("gather_records: videos missing or duplicated across ranks")
    keys = ("rgb_img_pred", "rgb_fea_comm", "op_img_pred", "op_fea_comm")
    return {k + "_records": [merged[v][k] for v in range(n_videos)] for k in keys}


def all_gather_scores(scores: torch.Tensor, group=None) -> torch.Tensor:
    """Per-frame score vector of equal length on every rank -> [world, n] (NCCL all_gather over NVLink; latency-bound)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scores.unsqueeze(0)
    out = torch.empty((dist.get_world_size(group),) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    dist.all_gather(list(out.unbind(0)), scores.contiguous(), group=group)
    return out


# --------------------------------------------------------------------------------------------------
# training: EMA statistics + gradients
# --------------------------------------------------------------------------------------------------
def install_stats_allreduce(group=None):
    """Make every memory bank update use GLOBAL-batch assignment statistics (single-GPU semantics of unet.py:298-309):
    counts[M] and embed_sum[D,M] of a module are flattened into one buffer and sum-all-reduced before the EMA step."""

    def hook(tensors):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n

    _modules._Hooks.stats_allreduce = hook
    return hook


def uninstall_stats_allreduce():
    _modules._Hooks.stats_allreduce = None


def install_sync_bn(model=None, group=None):
    """GLOBAL-batch BatchNorm statistics under data parallelism -- the semantics of the reference's single-GPU step
    (BatchNorm inside `double_conv`, unet.py:11-16; step structure train_helper.py:291-343):

    * the AMFT block's own kernels: the per-channel sums of the forward statistics and of the BatchNorm backward are
      sum-all-reduced between the two stages of ammc_bn_batch_stats_staged / ammc_bn_backward_staged;
    * the stock torch.nn.BatchNorm2d layers of the U-Net encoder / decoder (left on cuDNN for training): converted to
      torch.nn.SyncBatchNorm when `model` is given.  Returns the (possibly converted) model.
    Without this call every rank normalises with its own batch (stated as per-rank BN in DESIGN.md)."""
    from . import functions as _F

    def allreduce(sums):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    _F.BN_SYNC["allreduce"] = allreduce
    if model is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        bridge = getattr(model, "bridge", None)
        # the bridge's BatchNorm2d modules are parameter containers of our own kernels: keep them as they are
        keep = {id(m) for m in bridge.modules()} if bridge is not None else set()

        def convert(mod):
            for name, child in list(mod.named_children()):
                if id(child) in keep:
                    continue
                if isinstance(child, torch.nn.BatchNorm2d):
                    setattr(mod, name, torch.nn.SyncBatchNorm.convert_sync_batchnorm(child, group))
                else:
                    convert(child)
            return mod
        model = convert(model)
    return model


def uninstall_sync_bn():
    from . import functions as _F
    _F.BN_SYNC["allreduce"] = None


def allreduce_gradients(params, group=None, average: bool = True):
    """One flat all-reduce of all gradients (25 M fp32 = 100 MB for the whole generator; sub-millisecond class on NVLink 5)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
