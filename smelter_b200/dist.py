"""Multi-GPU host plumbing: one process per GPU, batches sharded across ranks.

The reference is single-device; this is new.  The path shards by independent images, so there is NO steady-state
collective: the only exchange is the one-time broadcast of the packed weight arena from rank 0 (NCCL over
NVLink, through the engine's own communicator — csrc/nccl_shim.cc).  torch.distributed is used for the
rendezvous, barriers and the max-over-ranks timing reduction only.
"""
from __future__ import annotations

import os
from typing import List, Tuple


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank`; remainders go to the lowest ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str):
    """torchrun-style rendezvous (MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE from the environment)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist


def share_bytes(payload: bytes, src: int = 0) -> bytes:
    """Every rank receives rank `src`'s byte string (used for the 128-byte NCCL unique id)."""
    import torch.distributed as dist

    box: List[object] = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return bytes(box[0])  # type: ignore[arg-type]


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_equal(value: int, device=None) -> bool:
    """True when every rank holds the same 64-bit value (weight-arena checksums after the broadcast)."""
    import torch
    import torch.distributed as dist

    lo = torch.tensor([value & 0xFFFFFFFF, value >> 32], dtype=torch.int64, device=device)
    mx, mn = lo.clone(), lo.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return bool((mx == mn).all().item())
