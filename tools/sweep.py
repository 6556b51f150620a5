"""BASELINE configs[4]: max_ac_errors 0..4 x injected BER sweep -- Gbit/s and detection rate, at 1 GPU
or (under torchrun) N GPUs.

    python tools/sweep.py [--symbols N] > profiles/r02_sweep_config5_n1.json
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py > profiles/r02_sweep_config5_n8.json

For every (k, BER) cell ONE synthetic stream of `--symbols` symbols (planted packet every 10 000,
BER injected into every symbol) is scanned promiscuously with max_ac_errors = k: on one GPU with
btbb_b200_find_ac_dev, on N GPUs cut into N contiguous shards through the C ABI's sharded scan
(btbb_b200_find_ac_sharded_*, hit records gathered on every rank).  The gathered hit list of the
WHOLE stream is compared, record for record, with the unmodified reference run over the same whole
stream on all host threads (oracle/_ref, ref_find_all_mt_hits; the oracle port where _ref was not
built) -- rank 0 does that once per cell.  Detection rate = planted access codes reported at their
exact offset with the right LAP / planted, over every planted packet, from the same list (so the GPU
and the reference curves coincide exactly when the lists do)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

BERS = [0.0, 0.001, 0.005, 0.01, 0.02, 0.05]
MIX = ("ID", "DM1", "DM3", "DH1", "FHS")
STRIDE = 10000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--symbols", type=int, default=2 * 10**9, help="symbols of the whole stream of a cell")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kmax", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-digests", default=os.path.join(ROOT, "profiles", "r02_sweep_reference_digests.json"),
                    help="count + sha256 of the reference's hit list per cell: written by a run that computes them "
                         "(the reference over the whole stream on the host), compared against by a run that finds them there")
    a = ap.parse_args()
    import torch, torch.distributed as dist
    import util
    from util import B
    from libbtbb_b200 import sharding
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = B.lib()
    n = a.symbols
    b, e = sharding.shard_range(n, rank, world)
    rb, rs = sharding.shard_read_span(n, rank, world, n + sharding.SEAM)
    d = torch.empty(rs - rb, dtype=torch.uint8, device="cuda")
    cap_all = n // STRIDE * 2 + (1 << 21)
    cap = cap_all if world == 1 else (e - b) // STRIDE * 2 + (1 << 20)
    d_all = torch.zeros((cap_all, 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    threads = os.cpu_count() or 1
    out = []
    # ground truth of every planted packet (offset, LAP), once
    truth = None
    if rank == 0:
        cfg0 = B.synth_cfg(n + sharding.SEAM, stride=STRIDE, mix=MIX)
        pl = B.Planted()
        n_slots = n // STRIDE
        truth = np.zeros((n_slots, 2), dtype=np.int64)
        for sl in range(n_slots):
            lib.btbb_b200_synth_planted(C.byref(cfg0), sl, C.byref(pl))
            truth[sl] = (pl.offset, pl.lap)
        truth = truth[truth[:, 0] + 64 <= n]
    host = None
    import hashlib
    recorded = {}
    if os.path.exists(a.cpu_digests):
        recorded = json.load(open(a.cpu_digests))
    fresh = {}
    for k in range(a.kmax + 1):
        ctx = B.Context(local, k)
        sh = sharding.ShardedScan(ctx, cap) if world > 1 else None
        cpu_lib = None
        need_cpu = any(f"{n}/{k}/{ber}" not in recorded for ber in BERS)
        if rank == 0 and not a.no_cpu and need_cpu:
            if util.have_ref():
                # the reference builds its syndrome map once per loaded library, for the first k > 0
                # (bluetooth_packet.c:288): every k gets its own copy of the library
                import shutil
                tmp = f"/tmp/libbtbb_ref_k{k}_{os.getpid()}.so"
                shutil.copy(util.REF_SO, tmp)
                cpu_lib = C.CDLL(tmp)
                os.remove(tmp)
                kind = "reference"
            else:
                cpu_lib, kind = util.oracle(), "port"
        for ber in BERS:
            cfg = B.synth_cfg(rs - rb, stride=STRIDE, ber=ber, mix=MIX, first_symbol=rb)
            B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), st)); torch.cuda.synchronize()

            def scan():
                if world == 1:
                    cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_all.data_ptr(), cap_all, k=k, stream=st)
                    assert rc == 0
                    return cnt
                counts, total, rc = sh.scan_all(d.data_ptr(), e - b, b, d_all.data_ptr(), cap_all, k=k, stream=st)
                assert rc == 0
                return total
            for _ in range(2):
                cnt = scan()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                cnt = scan()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            if rank == 0:
                hits = d_all[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
                i = np.searchsorted(hits["offset"], truth[:, 0])
                i = np.minimum(i, max(len(hits) - 1, 0))
                found = int(((hits["offset"][i] == truth[:, 0]) & (hits["lap"][i] == truth[:, 1])).sum()) if len(hits) else 0
                cell = {"k": k, "ber": ber, "n_gpus": world, "symbols": n, "gbit_s": n / (ms / 1e3) / 1e9, "ms": ms, "hits": int(cnt),
                        "planted": int(len(truth)), "detection_rate": found / len(truth)}
                key = f"{n}/{k}/{ber}"
                sha = hashlib.sha256(hits.tobytes()).hexdigest()
                if key in recorded and not a.no_cpu:
                    r = recorded[key]
                    cell.update({"matches_cpu": bool(r["hits"] == cnt and r["sha256"] == sha), "cpu_kind": r["kind"], "cpu_hits": r["hits"],
                                 "cpu_gbit_s": r.get("gbit_s"), "cpu_threads": r.get("threads"),
                                 "compared": "count + sha256 of every hit record of the whole stream, against the digest the reference "
                                             "produced for this cell (" + os.path.basename(a.cpu_digests) + ")"})
                elif cpu_lib is not None:
                    # the whole stream on the host, through the reference
                    if host is None:
                        host = np.empty(n + sharding.SEAM, dtype=np.uint8)
                    cfgh = B.synth_cfg(n + sharding.SEAM, stride=STRIDE, ber=ber, mix=MIX)
                    assert util.oracle().orc_synth(C.byref(cfgh), host.ctypes.data, threads) == 0
                    want = np.zeros(cap_all, dtype=B.HIT_DTYPE)
                    t0 = time.perf_counter()
                    if kind == "reference":
                        cpu_lib.btbb_init(k)
                        cpu_lib.ref_find_all_mt_hits.restype = C.c_int64
                        cpu_lib.ref_find_all_mt_hits.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int64]
                        nw = cpu_lib.ref_find_all_mt_hits(host.ctypes.data, n, B.LAP_ANY, k, threads, want.ctypes.data, cap_all)
                    else:
                        cpu_lib.orc_init(k)
                        nw = cpu_lib.orc_find_all(host.ctypes.data, n, B.LAP_ANY, k, want.ctypes.data, cap_all)
                    cpu_s = time.perf_counter() - t0
                    cell.update({"matches_cpu": bool(nw == cnt and want[:nw].tobytes() == hits.tobytes()), "cpu_kind": kind,
                                 "cpu_hits": int(nw), "cpu_gbit_s": n / cpu_s / 1e9, "cpu_threads": threads,
                                 "compared": "every hit record of the whole stream"})
                    fresh[key] = {"hits": int(nw), "sha256": hashlib.sha256(want[:nw].tobytes()).hexdigest(), "kind": kind,
                                  "gbit_s": n / cpu_s / 1e9, "threads": threads}
                out.append(cell)
                print(json.dumps(cell), flush=True)
        if sh:
            sh.close()
        ctx.close()
    if rank == 0 and fresh:
        recorded.update(fresh)
        json.dump(recorded, open(a.cpu_digests, "w"), indent=0, sort_keys=True)
    if rank == 0:
        print(json.dumps({"summary": "all_match", "value": all(o.get("matches_cpu", True) for o in out), "cells": len(out)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
