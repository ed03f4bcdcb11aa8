"""CPU, world_size 2 over gloo: the N>1 host logic (video sharding, record gather, EMA-statistic and gradient all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ammcnet_aaai2021_b200 import dist as adist
from ammcnet_aaai2021_b200 import modules as amod

AVENUE = [1439, 1211, 923, 947, 1007, 1283, 605, 36, 1175, 841, 472, 1271, 549, 507, 1001, 740, 426, 294, 248, 273, 76]


def test_lpt_partition_balances_and_covers():
    for world in (1, 2, 4, 8):
        parts = adist.lpt_partition(AVENUE, world)
        assert sorted(i for p in parts for i in p) == list(range(len(AVENUE)))
        loads = [sum(AVENUE[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(AVENUE)
        assert max(loads) <= 1.15 * sum(AVENUE) / world + 1
    assert adist.lpt_partition([], 2) == [[], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lengths = [23, 40, 9, 17, 31]
        mine = adist.lpt_partition(lengths, world)[rank]
        local = {}
        for v in mine:
            base = np.arange(lengths[v], dtype=np.float32) + 100 * v
            local[v] = {"rgb_img_pred": base, "rgb_fea_comm": base + 0.5, "op_img_pred": base + 1, "op_fea_comm": base + 2}
        res = adist.gather_records(local, len(lengths))
        if rank == 0:
            assert [len(a) for a in res["rgb_img_pred_records"]] == lengths
            for v in range(len(lengths)):
                assert res["rgb_fea_comm_records"][v][0] == 100 * v + 0.5
        else:
            assert res is None
        # EMA statistics: every rank must end up with the global sums
        adist.install_stats_allreduce()
        counts = torch.full((7,), float(rank + 1))
        esum = torch.arange(21, dtype=torch.float32).view(3, 7) * (rank + 1)
        amod._Hooks.stats_allreduce([counts, esum])
        tot = sum(r + 1 for r in range(world))
        assert torch.equal(counts, torch.full((7,), float(tot)))
        assert torch.equal(esum, torch.arange(21, dtype=torch.float32).view(3, 7) * tot)
        adist.uninstall_stats_allreduce()
        # gradients: averaged over ranks
        lin = torch.nn.Linear(4, 3)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank))
        adist.allreduce_gradients(list(lin.parameters()))
        for p in lin.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, (world - 1) / 2.0))
        # adversarial step (tools/train_step.py --gan): the discriminator is a second parameter group with its own all-reduce;
        # parameters without a gradient (the frozen flow estimator, a discriminator frozen for the generator update) are skipped
        import ammcnet_aaai2021_b200 as A0
        D = A0.PixelDiscriminator(3, [4, 8, 8, 8], use_norm=False)
        d_params = list(D.parameters())
        for i, p in enumerate(d_params):
            p.grad = torch.full_like(p, float(rank * (i + 1))) if i % 2 == 0 else None
        adist.allreduce_gradients(d_params)
        for i, p in enumerate(d_params):
            if i % 2 == 0:
                assert torch.allclose(p.grad, torch.full_like(p, (i + 1) * (world - 1) / 2.0))
            else:
                assert p.grad is None
        # global-batch BatchNorm: the hook sums the per-channel statistics over ranks and reports the world size; stock
        # BatchNorm2d layers outside the AMFT block become SyncBatchNorm, the block's own parameter containers stay
        import ammcnet_aaai2021_b200 as A
        from ammcnet_aaai2021_b200 import functions as F_
        net = torch.nn.Module()
        net.enc = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
        net.bridge = A.bridge(in_c=64)
        net = adist.install_sync_bn(net)
        assert isinstance(net.enc[1], torch.nn.SyncBatchNorm) and type(net.bridge.O2F.conv[1]) is torch.nn.BatchNorm2d
        sums = torch.arange(6, dtype=torch.float64) * (rank + 1)
        assert F_.BN_SYNC["allreduce"](sums) == world
        assert torch.equal(sums, torch.arange(6, dtype=torch.float64) * tot)
        adist.uninstall_sync_bn()
        assert F_.BN_SYNC["allreduce"] is None
        sc = adist.all_gather_scores(torch.full((2, 5), float(rank)))
        assert sc.shape == (world, 2, 5) and float(sc[1].mean()) == 1.0
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
