"""Why is the host-buffer path slow?  Times btbb_b200_find_ac_host on pinned / pageable input."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
_os_pageable = os.environ.get("PROBE_PAGEABLE", "0") == "1"
from libbtbb_b200 import binding as B
lib = B.lib()
n = int(float(os.environ.get("PROBE_SYMBOLS", "2e9")))
cfg = B.synth_cfg(n + 63, stride=10000)
d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()
hp = torch.empty(n + 63, dtype=torch.uint8, pin_memory=True); hp.copy_(d); torch.cuda.synchronize()
hn = hp.numpy().copy() if _os_pageable else None
ctx = B.Context(0, 2)
if os.environ.get("PROBE_TRACE") == "1":
    ctx.set_option(B.OPT_TRACE, 1)
cap = n // 10000 + (1 << 20)
hits = np.zeros(cap, dtype=B.HIT_DTYPE); got = C.c_int64(0)
import os as _os
print("host threads", _os.cpu_count())
for nt in _os.environ.get("PROBE_THREADS", "0").split(","):
 for ns in _os.environ.get("PROBE_STREAMS", "1").split(","):
  ctx.set_option(B.OPT_PACK_THREADS, int(nt)); ctx.set_option(B.OPT_PACK_STREAMS, int(ns))
  for name, ptr in (("pinned", hp.data_ptr()),) + ((("pageable", hn.ctypes.data),) if hn is not None else ()):
    for it in range(4):
        t = time.perf_counter()
        B.check(lib.btbb_b200_find_ac_host(ctx.h, ptr, n, B.LAP_ANY, 2, hits.ctypes.data, cap, C.byref(got)))
        dt = time.perf_counter() - t
        print("threads", nt, "streams", ns, name, it, f"{dt*1e3:.1f} ms  {n/dt/1e9:.1f} GB/s  hits {got.value}")
