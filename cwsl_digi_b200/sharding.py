"""Receiver-to-rank partition of the station (one receiver/band per GPU) and the
station-level gather of per-slot audio. The data path itself needs no collective: receivers share
nothing (source/CWSL_DIGI.cpp:80, 115-129). torch.distributed is plumbing only."""
from __future__ import annotations

from typing import List, Sequence


def receivers_of_rank(n_receivers: int, rank: int, world: int) -> List[int]:
    """Round-robin: receiver r is demodulated by rank r mod world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_receivers, world))


def owner_of_receiver(receiver: int, world: int) -> int:
    return receiver % world


def gather_slot_audio(local: Sequence, dst: int = 0):
    """Gather one equal-sized int16 tensor per rank to ``dst`` (NCCL on GPUs, gloo on CPU).
    ``local`` is this rank's tensor; returns the list on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    # int16 is not a collective dtype of either backend: ship the bytes
    wire = local.contiguous().view(torch.uint8)
    bucket = [torch.empty_like(wire) for _ in range(world)] if rank == dst else None
    dist.gather(wire, bucket, dst=dst)
    return [b.view(local.dtype) for b in bucket] if bucket is not None else None


def max_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
