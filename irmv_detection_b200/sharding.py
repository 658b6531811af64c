"""Multi-GPU sharding of frame batches: one process per GPU, no data-path collective.

The path has no exchange step (the reference runs three independent
engines on three streams, /root/reference/src/irm_detector.cpp:35-38), so frames are split
contiguously across ranks and each rank runs its own engine.  torch.distributed is used only
for the barrier and the max-over-ranks timing reduction of the bench.
"""
from __future__ import annotations

import os
from typing import Tuple


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units owned by `rank`; remainders go to low ranks."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_process_group(backend: str):
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_max_sum(value_max: float, value_sum: float, device: str = "cpu") -> Tuple[float, float]:
    """(max over ranks of value_max, sum over ranks of value_sum)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value_max, value_sum
    a = torch.tensor([value_max], dtype=torch.float64, device=device)
    b = torch.tensor([value_sum], dtype=torch.float64, device=device)
    dist.all_reduce(a, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(a.item()), float(b.item())
