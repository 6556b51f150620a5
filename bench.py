#!/usr/bin/env python
"""bench.py -- Gbit/s of input symbols through btbb_find_ac (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE configs[1] -- promiscuous LAP discovery
(lap = LAP_ANY, max_ac_errors = 2) over a 10 Gbit synthetic symbol stream, one byte per
symbol, per GPU.  With N > 1 the global stream is N x 10 Gbit cut into contiguous shards
(72-symbol seam overlap), one rank per GPU, weak scaling; the per-rank sorted hit lists are
gathered with one padded NCCL all-gather.  A step = one complete pass: scan kernel, hit
count read-back, ordering pass (and the gather when N > 1), input resident in HBM.

--impl reference times the reference's own CPU code (oracle/_ref when it was built, else
the oracle port) on all host threads over a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbit/s of input symbols through btbb_find_ac"
UNIT = "Gbit/s"
SYMBOLS_PER_GPU = 10**10
K_ERRORS = 2
STRIDE = 10000
MIX = ("ID", "DM1", "DM3", "DH1", "FHS")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_quota():
    """CPUs the container may use at once (cgroup v2 cpu.max), or None when unlimited/unknown."""
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if q == "max" else round(int(q) / int(p), 2)
    except Exception:
        return None


def cpu_scan_rate(sample_syms, min_seconds, threads, k):
    """Time the reference C path (or the oracle port) over a host-resident sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import util
    from libbtbb_b200 import binding as B
    kind = "reference" if util.have_ref() else "port"
    cfg = B.synth_cfg(sample_syms + 63, stride=STRIDE, mix=MIX)
    buf = None
    try:
        import torch
        if torch.cuda.is_available():   # generate the input quickly on the device, off the clock
            d = torch.empty(sample_syms + 63, dtype=torch.uint8, device="cuda")
            B.check(B.lib().btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0))
            buf = d.cpu().numpy()
            del d
    except Exception:
        buf = None
    if buf is None:
        buf = B.synth_host(cfg)
    if kind == "reference":
        L = util.ref()
        assert L.btbb_init(k) == 0
        fn = L.ref_find_all_mt
    else:
        L = util.oracle()
        assert L.orc_init(k) == 0
        fn = L.orc_find_all_mt
    hits = C.c_int64(0)
    fn(buf.ctypes.data, min(sample_syms, 1 << 24), B.LAP_ANY, k, threads, C.byref(hits))   # warm-up
    total_t, total_s, reps = 0.0, 0, 0
    while (reps == 0 or total_t < min_seconds) and reps < 64:
        total_t += fn(buf.ctypes.data, sample_syms, B.LAP_ANY, k, threads, C.byref(hits))
        total_s += sample_syms
        reps += 1
    return {"value": total_s / total_t / 1e9, "unit": UNIT, "cores": threads, "cpu_quota": cpu_quota(), "kind": kind,
            "sample": f"{reps} x {sample_syms} symbols of the same synthetic stream (stride {STRIDE}, BER 0), "
                      f"promiscuous k={k}, contiguous chunks over {threads} pthreads, gcc -O2",
            "hits_per_pass": int(hits.value), "seconds": round(total_t, 3)}, total_t, reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 1 << 30
    per_step = []
    base = None
    for i in range(args.warmup + args.steps):
        base, t, reps = cpu_scan_rate(sample, 0.0, threads, K_ERRORS)
        if i >= args.warmup:
            per_step.append(t / max(reps, 1))
    ms = 1e3 * sum(per_step) / len(per_step)
    value = sample / (ms / 1e3) / 1e9
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "promiscuous btbb_find_ac (LAP_ANY, max_ac_errors=2), 10 Gbit synthetic stream; "
                                   "each step = one 2^30-symbol sample of it on the host cores",
                       "symbols_per_step": sample, "max_ac_errors": K_ERRORS},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the one JSON line: everything else that writes to fd 1 (NCCL's version
    banner, library chatter) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--symbols", type=int, default=SYMBOLS_PER_GPU, help="symbols per GPU (default 10^10)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from libbtbb_b200 import binding as B
    from libbtbb_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: launch N > 1 with python -m torch.distributed.run --nproc-per-node N ...")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = B.lib()
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)

    # ---- this rank's shard of the global stream: positions [begin, end) + seam ----
    per = args.symbols
    free, _ = torch.cuda.mem_get_info()
    if per + (1 << 30) > free:
        per = int((free - (1 << 30)) // 2**20 * 2**20)
    total_positions = per * world
    begin, end = sharding.shard_range(total_positions, rank, world)
    rb, rs = sharding.shard_read_span(total_positions, rank, world, total_positions + sharding.SEAM)
    n = end - begin
    cfg = B.synth_cfg(rs - rb, stride=STRIDE, mix=MIX, first_symbol=rb)
    d_stream = torch.empty(rs - rb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d_stream.data_ptr(), st))
    torch.cuda.synchronize()
    cap = n // STRIDE + (1 << 20)
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    d_count = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx = B.Context(local, K_ERRORS)

    # hit buffers alternate between steps: the exchange of step i is still in flight (copy engines)
    # while step i + 1 scans
    d_hits2 = torch.zeros_like(d_hits) if world > 1 else None
    peer = None
    if world > 1 and os.environ.get("BTBB_B200_GATHER", "peer") == "peer":
        try:
            peer = sharding.PeerGather(cap, torch.device("cuda", local))
        except Exception as ex:                     # no symmetric memory on this box / build: NCCL gather
            sys.stderr.write(f"[bench] peer-memory gather unavailable ({ex!r}); using the NCCL all-gather\n")
            peer = None
    flip = [0]
    if world > 1:
        ctx.set_offset_bias(begin)                     # the kernels report global offsets

    def step():
        buf = d_hits if (world == 1 or flip[0] == 0) else d_hits2
        flip[0] ^= 1
        cnt, rc = ctx.find_ac_dev(d_stream.data_ptr(), n, buf.data_ptr(), cap, lap=B.LAP_ANY, k=K_ERRORS, stream=st)
        assert rc == 0
        if world > 1:
            _, counts = sharding.gather_hits(buf[:cnt], concat=False)
            return cnt, int(sum(counts))
        return cnt, cnt

    def run_steps(k):
        """k whole steps.  With the peer-memory exchange the steps are pipelined: the scan of step
        i + 1 is enqueued before the exchange of step i, whose copies then run on the copy engines
        underneath it; the last exchange is complete when this returns."""
        if peer is None:
            for _ in range(k):
                local_hits, total_hits = step()
            return local_hits, total_hits
        bufs = (d_hits, d_hits2)
        # PeerGather's generations alternate like the hit buffers: generation == buffer index
        assert peer.gen == flip[0]
        peer.wait_sent(flip[0])
        ctx.find_ac_dev_begin(d_stream.data_ptr(), n, bufs[flip[0]].data_ptr(), cap, lap=B.LAP_ANY, k=K_ERRORS, stream=st)
        for i in range(k):
            cur = bufs[flip[0]]
            flip[0] ^= 1
            cnt, rc = ctx.find_ac_dev_end()
            assert rc == 0
            if i + 1 < k:
                peer.wait_sent(flip[0])
                ctx.find_ac_dev_begin(d_stream.data_ptr(), n, bufs[flip[0]].data_ptr(), cap, lap=B.LAP_ANY, k=K_ERRORS, stream=st)
            peer.start(cur[:cnt])
        _, counts = peer.finish()
        return cnt, int(sum(counts))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    local_hits, total_hits = run_steps(warmup)
    if peer is not None:
        # the peer-memory exchange must deliver exactly what the NCCL all-gather does
        buf = d_hits2 if flip[0] == 0 else d_hits
        slots, counts = peer.finish()
        ref_slots, ref_counts = sharding.gather_hits(buf[:local_hits], concat=False)
        assert counts == ref_counts, (counts, ref_counts)
        for r, c in enumerate(counts):
            assert torch.equal(slots[r, :c], ref_slots[r, :c]), f"peer gather differs from NCCL in slot {r}"
    # ---- timed region: K whole steps, device clock, max over ranks ----
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    local_hits, total_hits = run_steps(steps)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    value = total_positions / (ms_step / 1e3) / 1e9
    # kernels of this library launched per step: bulk scan, tile kernel on the ragged tail,
    # slab_scan + slab_sort (ordering)
    launches_per_step = 4

    # ---- roofline of the dominant kernel: the scan alone, CUDA events on its stream ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        B.check(lib.btbb_b200_find_ac_enqueue(ctx.h, d_stream.data_ptr(), n, B.LAP_ANY, K_ERRORS, d_hits.data_ptr(),
                                              cap, d_count.data_ptr(), st))
        b.record()
    torch.cuda.synchronize()
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    peak, peak_src = measured_peak_gbs()
    alg_bytes = n + 63 + 16 * local_hits          # 1 B read per position + 16 B per hit
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_promisc_v7<0,5,1> (bulk, scan_v7.cuh) + tile kernel on the ragged tail", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms}
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            tr = json.load(open(tf))
            roofline["traffic"] = tr.get("dram_bytes_per_launch")
            roofline["traffic_note"] = tr.get("note")
        except Exception:
            pass

    # ---- e2e: the same scan through the host-buffer C-ABI call, pinned host input ----
    e2e = None
    if not args.no_e2e:
        try:
            h_stream = torch.empty(n + 63, dtype=torch.uint8, pin_memory=True)
            h_stream.copy_(d_stream[: n + 63])
            torch.cuda.synchronize()
            h_hits = np.zeros(cap, dtype=B.HIT_DTYPE)
            got = C.c_int64(0)
            e_steps = min(steps, 3)
            # one GPU per host: the library packs 32 symbols/word on the host cores before the copy
            # (PCIe-bound otherwise).  With several ranks on one host the cores are shared while
            # every GPU has its own PCIe link, so the byte-format copy is the faster route there.
            packed_route = world == 1 and os.environ.get("BTBB_B200_HOST") != "bytes"
            if not packed_route:
                os.environ["BTBB_B200_HOST"] = "bytes"

            def e_step():
                B.check(lib.btbb_b200_find_ac_host(ctx.h, h_stream.data_ptr(), n, B.LAP_ANY, K_ERRORS,
                                                   h_hits.ctypes.data, cap, C.byref(got)))
            e_step()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); d_stream[: n + 63].copy_(h_stream, non_blocking=True); c1.record()
            torch.cuda.synchronize()
            h2d_gbs = (n + 63) / (c0.elapsed_time(c1) / 1e3) / 1e9
            t0 = time.perf_counter()
            for _ in range(e_steps):
                e_step()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            assert got.value == local_hits
            e2e = {"value": total_positions / (float(dt.item()) / e_steps) / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": 4 * ((n + 63 + 31) // 32) if packed_route else n + 63,
                   "d2h_bytes_per_step": 16 * int(got.value) + 8, "steps": e_steps,
                   "host_input_bytes_per_step": n + 63,
                   "route": ("host cores pack 32 symbols/word (SSE2 movemask, 16 threads) -> H2D of packed words -> "
                             "packed bulk kernel" if packed_route else "H2D of the byte stream in 64 MiB chunks on two streams"),
                   "api": "btbb_b200_find_ac_host (pinned host stream -> sorted host hit records)",
                   "timer": "host wall clock around the blocking calls, max over ranks",
                   "plain_h2d_copy_GBps_same_buffer": round(h2d_gbs, 1)}
            del h_stream
        except Exception as ex:   # e.g. not enough pinnable host memory for 10 GB
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        del d_hits
        cpu, _, _ = cpu_scan_rate(1 << 30, 12.0, os.cpu_count() or 1, K_ERRORS)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "promiscuous btbb_find_ac (LAP_ANY, max_ac_errors=2) over a 10 Gbit synthetic "
                                       "symbol stream per GPU, 1 byte/symbol (BASELINE configs[1])",
                           "symbols_per_gpu": n, "total_symbols": total_positions, "max_ac_errors": K_ERRORS,
                           "planted_stride": STRIDE, "hits_total": total_hits, "seam_symbols": sharding.SEAM,
                           "l2": "input (10 GB/GPU) is far larger than the 126 MB L2; no flush needed",
                           "parallelism": f"contiguous shards x{world}, " + ("all-gather of hit records over NVLink peer memory (copy engines, overlapped with the next scan)" if peer is not None else "one NCCL all-gather of fixed-size hit-record slots")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * steps,
                "clocks": clocks}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
