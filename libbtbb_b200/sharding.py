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


_bufs = {}


def gather_hits(local_hits, group=None, concat=True):
    """allgatherv of 16-byte hit records.

    local_hits: uint8 tensor [n_local, 16] (device tensor with NCCL, CPU tensor with gloo)
    whose offsets are already global.  Returns (all_hits [n_total, 16], counts list); with
    concat=False the first item is the padded [world, cap, 16] receive buffer instead (rank
    r's records are buf[r, :counts[r]]), which skips one device copy.
    NCCL has no native allgatherv: one all_gather of the counts, then one all_gather of the
    records padded to the largest count.  Buffers are cached between calls.
    """
    world = dist.get_world_size(group)
    dev = local_hits.device
    n_local = torch.tensor([local_hits.shape[0]], dtype=torch.int64, device=dev)
    all_n = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_n, n_local, group=group)
    counts = all_n.tolist()                        # the one host sync of the exchange
    cap = max(max(counts), 1)
    cap = (cap + 65535) // 65536 * 65536           # round up so that the buffers get reused
    key = (str(dev), world, cap)
    if key not in _bufs:
        _bufs.clear()
        _bufs[key] = (torch.zeros((cap, 16), dtype=torch.uint8, device=dev),
                      torch.empty((world, cap, 16), dtype=torch.uint8, device=dev))
    send, recv = _bufs[key]
    send[: local_hits.shape[0]].copy_(local_hits)
    dist.all_gather_into_tensor(recv.view(world * cap, 16), send, group=group)
    if not concat:
        return recv, counts
    out = torch.cat([recv[r, :c] for r, c in enumerate(counts)], dim=0)
    return out, counts
