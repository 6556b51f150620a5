"""Developer tool: per-warp counters of the bulk promiscuous kernel (BTBB_B200_DBG=1).

    python tools/warp_dbg.py [--symbols N] [--mode v6]

Prints the distribution of leftover-loop trips and cycles per warp, to spot stragglers."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libbtbb_b200 import binding as B

ap = argparse.ArgumentParser()
ap.add_argument("--symbols", type=int, default=4 * 10**9)
ap.add_argument("--shift", type=int, default=0, help="start the scan this many symbols into the buffer")
ap.add_argument("--mode", default="v6", help="comma-separated BTBB_B200_SCAN values")
args = ap.parse_args()
ap2 = None
os.environ["BTBB_B200_DBG"] = os.environ.get("BTBB_B200_DBG", "1")
lib = B.lib()
n = args.symbols
cfg = B.synth_cfg(n + 127, stride=10000, mix=("ID", "DM1", "DM3", "DH1", "FHS"))
d = torch.empty(n + 127 + args.shift, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr() + args.shift, 0))
cap = n // 10000 + (1 << 20)
hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
ctx = B.Context(0, 2)
st = torch.cuda.current_stream().cuda_stream
for mode in args.mode.split(","):
  os.environ["BTBB_B200_SCAN"] = mode
  print("==", mode)
  for _ in range(2):
    B.check(lib.btbb_b200_find_ac_enqueue(ctx.h, d.data_ptr() + args.shift, n, B.LAP_ANY, 2, hits.data_ptr(), cap, cnt.data_ptr(), st))
  torch.cuda.synchronize()
  out = np.zeros((148 * 32, 8), dtype=np.uint32)
  lib.bt_dbg_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
  nw = lib.bt_dbg_read(ctx.h, out.ctypes.data, out.shape[0])
  o = out[:nw].astype(np.int64)
  print("warps", nw, "strips/warp", o[:, 3].min(), o[:, 3].max())
  for name, col in (("hot cyc (incl load)", 0), ("cold cyc", 1), ("cycles", 2), ("parks", 4), ("worst strip cyc", 5), ("load+pack cyc", 7)):
      v = o[:, col]
      print(f"{name:15s} min {v.min()} median {int(np.median(v))} p99 {int(np.percentile(v, 99))} max {v.max()} argmax {int(v.argmax())}")
  top = np.argsort(-o[:, 2])[:8]
  for w in top:
      print("warp", int(w), "cta", int(w) // (nw // 148), "Msym", round(int(w) * n / nw / 1e6, 1), o[w].tolist())
