"""Receiver-to-rank partition of the station (one receiver/band per GPU) and the
station-level gather of per-slot audio. The data path itself needs no collective: receivers share
nothing (source/CWSL_DIGI.cpp:80, 115-129). torch.distributed is plumbing only."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def receivers_of_rank(n_receivers: int, rank: int, world: int) -> List[int]:
    """Round-robin: receiver r is demodulated by rank r mod world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_receivers, world))


def owner_of_receiver(receiver: int, world: int) -> int:
    return receiver % world


def channel_slices_of_rank(n_receivers: int, n_channels: int, rank: int, world: int,
                           granule: int = 32) -> List[Tuple[int, int, int]]:
    """Secondary partition for a station with fewer receivers than GPUs: the ranks are
    dealt to the receivers round-robin and the ranks of one receiver split its channel list into contiguous
    slices (multiples of ``granule`` channels, the fast kernel's channel group, except the last). Every rank
    of a receiver is fed the same IQ; channels are independent (one SSBD each, source/Instance.cpp:186), so
    the concatenated slices equal the unsplit result bit for bit. With world <= n_receivers this degrades
    to receivers_of_rank with full slices. Returns [(receiver, ch_begin, ch_end), ...]."""
    if world <= 0 or not (0 <= rank < world) or granule <= 0:
        raise ValueError("bad rank/world")
    if n_receivers <= 0 or n_channels <= 0:
        return []
    if world <= n_receivers:
        return [(r, 0, n_channels) for r in receivers_of_rank(n_receivers, rank, world)]
    receiver = rank % n_receivers
    peers = list(range(receiver, world, n_receivers))         # ranks serving this receiver
    granules = -(-n_channels // granule)
    k = peers.index(rank)
    g0 = granules * k // len(peers)
    g1 = granules * (k + 1) // len(peers)
    lo, hi = min(g0 * granule, n_channels), min(g1 * granule, n_channels)
    return [(receiver, lo, hi)] if hi > lo else []


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
