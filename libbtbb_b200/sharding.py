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


class PeerGather:
    """The same all-gather of hit records, over NVLink peer memory with the COPY ENGINES.

    The scan kernel owns every SM (one persistent CTA per SM, all registers and shared memory),
    so an NCCL collective can only run after it, and the gather shows up in the step time.  Here
    every rank owns a symmetric-memory buffer of `world` slots (x 2 generations); a rank pushes its
    records -- a header record carrying the count, then the records -- into its slot on every
    peer with plain device-to-device copies on a side stream.  Those are DMA transfers: they
    overlap the NEXT step's scan completely.  No kernel is involved, so nothing waits for an SM.

    start(local_hits) enqueues the exchange of one step; finish() completes the most recent one on
    every rank (host-side wait + one NCCL barrier) and returns (slots [world, cap, 16], counts).
    """

    def __init__(self, cap_records, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.cap = int(cap_records)
        self.dev = device
        shape = (2, self.world, self.cap + 1, 16)
        self.buf = symm.empty(shape, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peers = [self.hdl.get_buffer(r, shape, torch.uint8) for r in range(self.world)]
        self.hdr = torch.zeros((2, 16), dtype=torch.uint8).pin_memory()      # header record per generation
        self.stream = torch.cuda.Stream(device=device)
        self.sent = [None, None]          # event per generation: its header has left pinned memory
        self.gen = 0
        self.last = None
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def start(self, local_hits):
        """local_hits: [n, 16] records, complete on the current stream (the caller has synchronised
        it or will not touch it until two more start() calls).  Everything enqueued here is a copy
        (host-to-device for the header, device-to-device for the records), so it does not need an
        SM while the next scan holds all of them."""
        n = int(local_hits.shape[0])
        if n > self.cap:
            raise ValueError("PeerGather: more records than the slots hold")
        g = self.gen
        self.gen ^= 1
        if self.sent[g] is not None:
            self.sent[g].synchronize()             # the header of two steps ago has been read
        self.hdr[g].view(torch.int64)[0] = n
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            for r in range(self.world):
                slot = self.peers[r][g, self.rank]
                slot[0].copy_(self.hdr[g], non_blocking=True)
                if n:
                    slot[1: n + 1].copy_(local_hits, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.sent[g] = ev
        self.last = g

    def wait_sent(self, g):
        """Host-side wait until generation g's copies have left this rank (its source buffer may
        then be overwritten)."""
        if self.sent[g] is not None:
            self.sent[g].synchronize()

    def finish(self):
        """Complete the most recent exchange on every rank: wait until this rank's copies have left,
        then meet the other ranks (one NCCL barrier) -- after it, every slot of the generation has
        landed everywhere.  Between start() calls nothing synchronises the ranks: slots of older
        generations may be overwritten by ranks that run ahead, only the latest one is read."""
        g = self.last
        self.stream.synchronize()
        dist.barrier(self.group)
        slots = self.buf[g]
        counts = slots.view(torch.int64)[:, 0, 0].tolist()
        return slots[:, 1:], counts


class ShardedScan:
    """The C ABI's multi-GPU scan (btbb_b200_shard_* / btbb_b200_find_ac_sharded_*, csrc/sharded.cu)
    driven from a torch.distributed program: torch only carries the 128-byte NCCL unique id from
    rank 0 to the others; the scan, the peer-memory exchange and the NCCL allgatherv form all live
    in the library."""

    def __init__(self, ctx, slot_records, group=None, nccl_only=False, copy_engines=False):
        import ctypes as C
        from . import binding as B
        self.B, self.C, self.ctx = B, C, ctx
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.world = dist.get_world_size(self.group) if self.group is not None else 1
        self.rank = dist.get_rank(self.group) if self.group is not None else 0
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            B.check(B.lib().btbb_b200_shard_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            t = ident.to(dev)
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else 0, group=self.group)
            ident = t.cpu()
        raw = (C.c_uint8 * 128)(*ident.tolist())
        B.check(B.lib().btbb_b200_shard_init(ctx.h, raw, self.rank, self.world, int(slot_records), (1 if nccl_only else 0) | (2 if copy_engines else 0)))
        self.copy_engines = bool(copy_engines)
        pm = C.c_int(0)
        B.check(B.lib().btbb_b200_shard_info(ctx.h, None, None, C.byref(pm)))
        self.peer_memory = bool(pm.value)
        self.slot = int(slot_records)

    def begin(self, d_ptr, search_length, first_position, lap=0xFFFFFFFF, k=2, stream=0):
        self.B.check(self.B.lib().btbb_b200_find_ac_sharded_begin(self.ctx.h, d_ptr, search_length, first_position, lap, k, stream))

    def end(self):
        n = self.C.c_int64(0)
        self.B.check(self.B.lib().btbb_b200_find_ac_sharded_end(self.ctx.h, self.C.byref(n)))
        return n.value

    def next(self, d_ptr, search_length, first_position, lap=0xFFFFFFFF, k=2, stream=0):
        """end of the pending scan + begin of the next one; returns the finished scan's hit count"""
        n = self.C.c_int64(0)
        self.B.check(self.B.lib().btbb_b200_find_ac_sharded_next(self.ctx.h, d_ptr, search_length, first_position, lap, k, stream,
                                                                 self.C.byref(n)))
        return n.value

    def gather(self):
        """(device pointer of slot 0, slot stride in records, counts list, total)"""
        C = self.C
        slots, stride, total = C.c_void_p(0), C.c_int64(0), C.c_int64(0)
        counts = (C.c_int64 * self.world)()
        self.B.check(self.B.lib().btbb_b200_find_ac_sharded_gather(self.ctx.h, C.byref(slots), C.byref(stride), counts, C.byref(total)))
        return slots.value, stride.value, list(counts), total.value

    def scan_all(self, d_ptr, search_length, first_position, d_all_ptr, max_all, lap=0xFFFFFFFF, k=2, stream=0):
        C = self.C
        counts = (C.c_int64 * self.world)()
        total = C.c_int64(0)
        rc = self.B.lib().btbb_b200_find_ac_sharded_dev(self.ctx.h, d_ptr, search_length, first_position, lap, k, d_all_ptr, max_all,
                                                       counts, C.byref(total), stream)
        self.B.check(rc, allow=(-4,))
        return list(counts), total.value, rc

    def close(self):
        self.B.check(self.B.lib().btbb_b200_shard_destroy(self.ctx.h))
