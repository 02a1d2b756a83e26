"""Multi-GPU plumbing for the stream-sharded path: one process per GPU, `torch.distributed` only.

The hot path has no exchange step — every stream owns its encoder / decoder / acquisition / filter state and never
reads another stream's data — so the only collective is ONE broadcast of the ~2.3 MB weight blob at
start-up; after that ranks never talk until the final timing reduction.
"""
import numpy as np


def stream_shard(n_streams_total, world, rank):
    """contiguous block of streams owned by `rank`: stream s -> rank s*world//S"""
    lo = (n_streams_total * rank) // world
    hi = (n_streams_total * (rank + 1)) // world
    return lo, hi


def broadcast_weights(dist, rank, blob, device="cpu"):
    """rank 0 holds `blob` (bytes); every rank returns the same bytes.  Works with nccl (device='cuda') and gloo."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    n = torch.tensor([len(blob) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, 0)
    if rank == 0:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def reduce_max(dist, value, device="cpu"):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
