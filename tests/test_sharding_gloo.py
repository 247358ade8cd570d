"""N>1 host logic on CPU: the reference's stride partition (misc/torchutils.py:66-68) and the label-map gather,
exercised with two gloo ranks."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from irn_b200.misc import torchutils
from oracle import steps


def test_split_matches_reference_partition():
    for n, k in [(10, 3), (64, 8), (5, 8), (10582, 8)]:
        mine = torchutils.split_indices(n, k)
        ref = steps.split_indices(n, k)
        assert all(np.array_equal(a, b) for a, b in zip(mine, ref))
        assert sorted(np.concatenate(mine).tolist()) == list(range(n))


def _rank(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 7
    mine = torchutils.split_indices(n, world)[rank]
    # each rank "labels" its images with their global index; ranks may own different counts -> pad, gather, unpad
    labels = torch.stack([torch.full((4, 4), int(i), dtype=torch.uint8) for i in mine]) if len(mine) else torch.empty((0, 4, 4), dtype=torch.uint8)
    cap = (n + world - 1) // world
    pad = torch.zeros((cap, 4, 4), dtype=torch.uint8)
    pad[:len(mine)] = labels
    got = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(got, pad)
    if rank == 0:
        merged = {}
        for r in range(world):
            for j, i in enumerate(torchutils.split_indices(n, world)[r]):
                merged[int(i)] = int(got[r][j, 0, 0])
        out.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert merged == {i: i for i in range(7)}
