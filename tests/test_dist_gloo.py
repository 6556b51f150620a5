"""world_size-2 gloo test of the multi-GPU host logic (sharding + hit gather) on CPU.  The
oracle stands in for the per-rank scan; the union must equal a single-rank scan."""
import os
import subprocess
import sys

import numpy as np

import util

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
import util
from util import B
from libbtbb_b200 import sharding
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), 2
O = util.oracle(); O.orc_init(2)
N = 700_001
cfg = B.synth_cfg(N + 63, stride=3000, ber=0.002, mix=('ID', 'DM1'))
b, e = sharding.shard_range(N, rank, world)
rb, rs = sharding.shard_read_span(N, rank, world, N + 63)
# each rank generates only its own shard (+ seam) of the capture
shard = B.synth_host(B.synth_cfg(rs - rb, stride=3000, ber=0.002, mix=('ID', 'DM1'), first_symbol=rb))
h = util.find_all(O, 'orc', shard, e - b, B.LAP_ANY, 2)
h['offset'] += b
local = torch.from_numpy(h.view(np.uint8).reshape(-1, 16).copy())
# a first, small exchange fixes the slot size; the real one then outgrows it on both ranks and
# has to re-agree (sharding.gather_hits)
few, few_counts = sharding.gather_hits(local[: 3 + rank])
assert few_counts == [3, 4] and few.shape[0] == 7
assert torch.equal(few[3:], local.new_tensor(few[3:]))
allh, counts = sharding.gather_hits(local)
if rank == 0:
    whole = B.synth_host(cfg)
    ref = util.find_all(O, 'orc', whole, N, B.LAP_ANY, 2)
    got = allh.numpy().reshape(-1).view(B.HIT_DTYPE)
    assert sum(counts) == len(ref) and counts[0] > 0 and counts[1] > 0, (counts, len(ref))
    assert got.tobytes() == ref.tobytes()
    print('OK', counts)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_shard_and_gather(product_lib, orc, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), util.ROOT, port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_shard_ranges_partition():
    from libbtbb_b200 import sharding
    for n in (0, 1, 7, 1000, 10**10):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
