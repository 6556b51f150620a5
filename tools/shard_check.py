"""Multi-GPU parity of the C ABI's sharded scan (run under torchrun, N >= 1):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_check.py

One global synthetic stream is cut into N contiguous shards; every rank scans its shard through
btbb_b200_find_ac_sharded_* (peer-memory exchange, then the NCCL allgatherv form) and the gathered
list must equal, record for record, what ONE GPU finds scanning the whole stream -- pipelined
(begin(i + 1) before end(i)) and not."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from libbtbb_b200 import binding as B
from libbtbb_b200 import sharding

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = B.lib()
total = int(os.environ.get("SHARD_CHECK_SYMBOLS", 400_000_000))
res = {"world": world}
for lap, k, stride in ((B.LAP_ANY, 2, 4000), (0x9E8B33, 1, 4000)):
    cfgw = B.synth_cfg(total + 72, stride=stride, ber=0.002, mix=("ID", "DM1", "DH1", "FHS"), n_laps=(1 if lap != B.LAP_ANY else 64))
    # the whole stream on this GPU (reference result), and this rank's shard on its own
    d_whole = torch.empty(total + 72, dtype=torch.uint8, device="cuda")
    B.check(lib.btbb_b200_synth_dev(C.byref(cfgw), d_whole.data_ptr(), 0))
    cap = total // stride * 2 + 4096
    d_ref = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    ctx = B.Context(local, 2)
    n_ref, rc = ctx.find_ac_dev(d_whole.data_ptr(), total, d_ref.data_ptr(), cap, lap=lap, k=k)
    assert rc == 0
    b, e = sharding.shard_range(total, rank, world)
    rb, rs = sharding.shard_read_span(total, rank, world, total + 72)
    cfgs = B.synth_cfg(rs - rb, stride=stride, ber=0.002, mix=("ID", "DM1", "DH1", "FHS"), n_laps=(1 if lap != B.LAP_ANY else 64), first_symbol=rb)
    d_shard = torch.empty(rs - rb, dtype=torch.uint8, device="cuda")
    B.check(lib.btbb_b200_synth_dev(C.byref(cfgs), d_shard.data_ptr(), 0))
    torch.cuda.synchronize()
    assert torch.equal(d_shard, d_whole[rb:rs])
    for nccl_only, ce in ((False, False), (False, True), (True, False)):
        sh = sharding.ShardedScan(ctx, cap, nccl_only=nccl_only, copy_engines=ce)
        d_all = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
        counts, n_all, rc = sh.scan_all(d_shard.data_ptr(), e - b, b, d_all.data_ptr(), cap, lap=lap, k=k)
        assert rc == 0 and n_all == n_ref, (n_all, n_ref)
        assert torch.equal(d_all[:n_all], d_ref[:n_ref]), "sharded list differs from the single-GPU list"
        # pipelined: three scans, scan i + 1 enqueued before scan i's records are pushed; only the
        # last exchange is gathered
        sh.begin(d_shard.data_ptr(), e - b, b, lap=lap, k=k)
        for i in range(2):
            n_loc = sh.next(d_shard.data_ptr(), e - b, b, lap=lap, k=k)
            assert n_loc == counts[rank]
        n_loc = sh.end()
        ptr, stride_r, counts2, n2 = sh.gather()
        assert counts2 == counts and n2 == n_ref
        # the library-owned slots hold the same records
        at = 0
        for r in range(world):
            class _Raw:      # a library-owned device buffer as a tensor
                __cuda_array_interface__ = {"shape": (max(counts2[r], 1), 16), "typestr": "|u1", "data": (ptr + r * stride_r * 16, False), "version": 2}
            slot = torch.as_tensor(_Raw(), device="cuda")[:counts2[r]]
            assert torch.equal(slot, d_ref[at:at + counts2[r]]), f"slot {r} differs"
            at += counts2[r]
        res[f"lap={lap:#x} k={k} nccl_only={nccl_only} copy_engines={ce}"] = {"hits": n_all, "counts": counts, "peer_memory": sh.peer_memory}
        sh.close()
    ctx.close()
    del d_whole, d_ref, d_shard, d_all
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps(res))
