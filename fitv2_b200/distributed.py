"""Multi-GPU sharding of the sampling job: one process per GPU, independent trajectories per rank
(sample_fitv2_ddp.py:51-57,230-239), no collective inside the step loop, one all-gather of the final
latents after the trajectory (template: sample_fit_ddp.py:185-186).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: str) -> Tuple[int, int, int]:
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def rank_seed(global_seed: int, world_size: int, rank: int) -> int:
    """sample_fitv2_ddp.py:54: seed = global_seed * world_size + rank."""
    return global_seed * world_size + rank


def draw_rank_inputs(global_seed: int, world_size: int, rank: int, n: int, tokens: int, channels: int, num_classes: int):
    """Noise and labels of one rank, both from the CPU generator seeded per rank (the script draws the noise on
    the CPU, :257-259; labels are drawn from the same CPU generator here so both arms see the same tensor)."""
    g = torch.Generator().manual_seed(rank_seed(global_seed, world_size, rank))
    z = torch.randn(n, tokens, channels, generator=g)
    y = torch.randint(0, num_classes, (n,), generator=g)
    return z, y


def shard_range(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard of `total` samples for `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_latents(z: torch.Tensor) -> torch.Tensor:
    """All-gather the final latents (n, N, C) of every rank -> (world*n, N, C), rank-major.  The only collective
    of the path; issued after the trajectory."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return z
    out = [torch.empty_like(z) for _ in range(dist.get_world_size())]
    dist.all_gather(out, z.contiguous())
    return torch.cat(out, dim=0)


def gather_images(images: torch.Tensor) -> torch.Tensor:
    """All-gather of the decoded uint8 images (n, H, W, 3) of every rank, the reference's own format for this step
    (sample_fit_ddp.py:185-186; commented out in sample_fitv2_ddp.py:328-329) -> (world*n, H, W, 3), rank-major."""
    return gather_latents(images)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
