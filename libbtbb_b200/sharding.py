"""Multi-GPU plumbing of the access-code scan: contiguous shards of the
symbol stream, one rank per GPU, and the single variable-length gather of hit records.

Window positions are independent, so the scan itself needs no inter-GPU traffic: rank r
searches positions [begin_r, end_r) and reads SEAM symbols past end_r (63 are necessary
for the 64-symbol window; the north star fixes the overlap at 72).  Ranges partition the
positions, so there are no duplicate hits and the rank-order concatenation of the
per-rank sorted lists is globally sorted.
"""
import torch
import torch.distributed as dist

SEAM = 72


def shard_range(n_positions, rank, world):
    """Positions [begin, end) searched by `rank` out of `world` contiguous shards."""
    return n_positions * rank // world, n_positions * (rank + 1) // world


def shard_read_span(n_positions, rank, world, total_symbols):
    """Symbols [begin, stop) rank must hold: its positions plus the seam overlap."""
    b, e = shard_range(n_positions, rank, world)
    return b, min(total_symbols, e + SEAM)


def gather_hits(local_hits, group=None):
    """allgatherv of 16-byte hit records.

    local_hits: uint8 tensor [n_local, 16] (device tensor with NCCL, CPU tensor with gloo)
    whose offsets are already global.  Returns (all_hits [n_total, 16], counts list).
    NCCL has no native allgatherv: one all_gather of the counts, then one all_gather of
    the records padded to the largest count.
    """
    world = dist.get_world_size(group)
    dev = local_hits.device
    n_local = torch.tensor([local_hits.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, 16), dtype=torch.uint8, device=dev)
    padded[: local_hits.shape[0]] = local_hits
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    out = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return out, counts
