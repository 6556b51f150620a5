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
_cap_hint = {}


def gather_hits(local_hits, group=None, concat=True):
    """allgatherv of 16-byte hit records.

    local_hits: uint8 tensor [n_local, 16] (device tensor with NCCL, CPU tensor with gloo)
    whose offsets are already global.  Returns (all_hits [n_total, 16], counts list); with
    concat=False the first item is the padded [world, cap, 16] receive buffer instead (rank
    r's records are buf[r, :counts[r]]), which skips one device copy.
    NCCL has no native allgatherv, so the exchange is ONE all_gather of fixed-size slots: slot r
    holds rank r's count (first 8 bytes of a 16-byte header record) followed by its records,
    padded to a capacity every rank derives the same way -- the largest count seen so far,
    rounded up.  The capacity is agreed once (one all_gather of the counts on the first call)
    and re-agreed only when some rank outgrows it; in steady state a call is one collective and
    one host read of the counts.  Buffers are cached between calls.
    """
    world = dist.get_world_size(group)
    dev = local_hits.device
    n = int(local_hits.shape[0])
    gkey = (str(dev), world, id(group))
    while True:
        cap = _cap_hint.get(gkey)
        if cap is None:
            n_local = torch.tensor([n], dtype=torch.int64, device=dev)
            all_n = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_n, n_local, group=group)
            cap = max(int(all_n.max().item()), 1)
            cap = (cap + cap // 64 + 4095) // 4096 * 4096      # 1.5 % head room, then a multiple of 4096
            _cap_hint[gkey] = cap
        key = (str(dev), world, cap)
        if key not in _bufs:
            _bufs.clear()
            _bufs[key] = (torch.zeros((cap + 1, 16), dtype=torch.uint8, device=dev),
                          torch.empty((world, cap + 1, 16), dtype=torch.uint8, device=dev))
        send, recv = _bufs[key]
        send.view(torch.int64)[0, 0] = n
        send[1: 1 + min(n, cap)].copy_(local_hits[:cap])
        dist.all_gather_into_tensor(recv.view(world * (cap + 1), 16), send, group=group)
        counts = recv.view(torch.int64)[:, 0, 0].tolist()      # the one host sync of the exchange
        if max(counts) <= cap:
            break
        del _cap_hint[gkey]                                    # some rank outgrew the slots: agree on a new size
    body = recv[:, 1:]
    if not concat:
        return body, counts
    out = torch.cat([body[r, :c] for r, c in enumerate(counts)], dim=0)
    return out, counts
