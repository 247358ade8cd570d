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


def _rank_writer(rank, world, port, out):
    """What bench.py / the steps do at N > 1, on CPU: every rank walks ITS stride shard through the chunked loader messages
    (step/_common.collate_chunk -> split_chunk), "labels" its images, and the label maps are GATHERED TO THE WRITER RANK only
    (dist.gather, the collective bench.py issues on a side stream) -- the other ranks receive nothing."""
    from torch.utils.data import DataLoader
    from irn_b200.step import _common
    from irn_b200.voc12 import dataloader as dl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    ds = dl.SyntheticMSF(n, size=(16, 24), scales=(1.0,), decode_only=True)
    shard = torchutils.split_dataset(ds, world)[rank]
    names, maps = [], []
    for msg in DataLoader(shard, shuffle=False, batch_size=4, num_workers=0, collate_fn=_common.collate_chunk):
        for pack in _common.split_chunk(msg):
            names.append(pack["name"][0])
            maps.append(torch.full((4, 4), int(pack["name"][0][-6:]), dtype=torch.uint8))     # the image's id as its "label map"
    cap = (n + world - 1) // world
    send = torch.zeros((cap, 4, 4), dtype=torch.uint8)
    if maps:
        send[:len(maps)] = torch.stack(maps)
    recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, gather_list=recv, dst=0)
    if rank == 0:
        merged = {}
        for r in range(world):
            for j, i in enumerate(torchutils.split_indices(n, world)[r]):
                merged[int(i)] = int(recv[r][j, 0, 0])
        out.put((merged, names))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_to_writer_over_chunked_shards_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_writer, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, names0 = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert merged == {i: i for i in range(11)}                                   # every image exactly once, at its own index
    assert names0 == ["2007_%06d" % i for i in range(0, 11, 2)]                  # rank 0 owns ids[0::2], in order (misc/torchutils.py:66-68)
