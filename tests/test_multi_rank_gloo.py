"""world_size-2 gloo test (CPU) of the N>1 host logic: per-rank seeds / shards and the final latents gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fitv2_b200.distributed import draw_rank_inputs, gather_images, gather_latents, max_over_ranks, rank_seed, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, y = draw_rank_inputs(0, world, rank, n=3, tokens=8, channels=16, num_classes=1000)
    z_final = z * 2 + rank                                  # stand-in for an independent trajectory
    allz = gather_latents(z_final)
    slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
    imgs = gather_images(torch.full((3, 4, 4, 3), 10 * rank + 1, dtype=torch.uint8))   # the reference's uint8 (n, H, W, 3) format
    ret[rank] = (allz.clone(), z_final.clone(), y.clone(), slow, imgs.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0].shape == (6, 8, 16) and torch.equal(r0[0], r1[0])              # every rank holds the same gather
    assert torch.equal(r0[0][:3], r0[1]) and torch.equal(r0[0][3:], r1[1])      # rank-major order
    assert not torch.equal(r0[1], r1[1]) and r0[3] == r1[3] == 11.0             # different seeds; MAX over ranks
    assert r0[4].dtype == torch.uint8 and r0[4].shape == (6, 4, 4, 3) and torch.equal(r0[4], r1[4])
    assert int(r0[4][0, 0, 0, 0]) == 1 and int(r0[4][3, 0, 0, 0]) == 11         # rank-major image order
    g = torch.Generator().manual_seed(rank_seed(0, 2, 1))
    assert torch.equal(torch.randn(3, 8, 16, generator=g) * 2 + 1, r1[1])       # seed = global*world + rank


def test_shard_ranges_cover_everything():
    for total, world in [(50000, 8), (7, 2), (64, 4), (5, 8)]:
        spans = [shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
