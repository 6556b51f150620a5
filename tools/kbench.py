"""Kernel bench / A-B check for the scan kernels (developer tool, not the contract bench).

    python tools/kbench.py [--symbols N] [--iters I] [--k K] [--check]

Times btbb_b200_find_ac_enqueue (scan only, CUDA events) on a device-resident synthetic
stream and, with --check, verifies the bulk kernel's sorted hit list against the tile
kernel (BTBB_B200_OPT_TILE_KERNEL_ONLY), which is itself pinned to the oracle by tests/."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libbtbb_b200 import binding as B

ap = argparse.ArgumentParser()
ap.add_argument("--symbols", type=int, default=4 * 10**9)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--k", type=int, default=2)
ap.add_argument("--check", action="store_true")
ap.add_argument("--misalign", type=int, default=0)
ap.add_argument("--lap", type=lambda x: int(x, 0), default=B.LAP_ANY, help="known LAP (default: promiscuous)")
args = ap.parse_args()

lib = B.lib()
n = args.symbols
cfg = B.synth_cfg(n + 63 + 64, stride=10000, mix=("ID", "DM1", "DM3", "DH1", "FHS"),
                   n_laps=(1 if args.lap != B.LAP_ANY else 64), fixed_lap=(args.lap & 0xffffff))
d = torch.empty(n + 63 + 64, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0))
torch.cuda.synchronize()
ptr = d.data_ptr() + args.misalign
cap = n // 10000 + (1 << 20)
hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
ctx = B.Context(0, args.k)
st = torch.cuda.current_stream().cuda_stream
out = {"symbols": n, "k": args.k, "lap": hex(args.lap)}
MODES = os.environ.get("KBENCH_MODES", "tile,bulk").split(",")      # tile kernels only / the shipped bulk kernels
for mode in MODES:
    ctx.set_option(B.OPT_TILE_KERNEL_ONLY, 1 if mode == "tile" else 0)
    for _ in range(3):
        B.check(lib.btbb_b200_find_ac_enqueue(ctx.h, ptr, n, args.lap, args.k, hits.data_ptr(), cap, cnt.data_ptr(), st))
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for a, b in ev:
        a.record()
        B.check(lib.btbb_b200_find_ac_enqueue(ctx.h, ptr, n, args.lap, args.k, hits.data_ptr(), cap, cnt.data_ptr(), st))
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    out[mode] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "GBps": n / (ms[len(ms) // 2] / 1e3) / 1e9,
                 "hits": int(cnt[0].item())}
    if args.check:
        c, rc = ctx.find_ac_dev(ptr, n, hits.data_ptr(), cap, lap=args.lap, k=args.k)
        out[mode]["sha"] = __import__("hashlib").sha256(hits[:c].cpu().numpy().tobytes()).hexdigest()[:16]
if args.check:
    out["match"] = all(out[m]["sha"] == out[MODES[0]]["sha"] and out[m]["hits"] == out[MODES[0]]["hits"] for m in MODES)
print(json.dumps(out))
