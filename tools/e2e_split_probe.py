"""btbb_b200_find_ac_host over 10^10 pinned symbols: share of the stream that travels as bytes (DMA, no CPU work)
while the host cores pack the rest (BTBB_B200_OPT_HOST_SPLIT_PERMILLE).  Results are flushed line by line."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libbtbb_b200 import binding as B
lib = B.lib()
n = int(float(os.environ.get("PROBE_SYMBOLS", "1e10")))
cfg = B.synth_cfg(n + 63, stride=10000)
d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()
hp = torch.empty(n + 63, dtype=torch.uint8, pin_memory=True); hp.copy_(d); torch.cuda.synchronize()
del d
ctx = B.Context(0, 2)
cap = n // 10000 + (1 << 20)
hits = np.zeros(cap, dtype=B.HIT_DTYPE); got = C.c_int64(0)
B.check(lib.btbb_b200_find_ac_host(ctx.h, hp.data_ptr(), n, B.LAP_ANY, 2, hits.ctypes.data, cap, C.byref(got)))
want = got.value
for split in [int(x) for x in os.environ.get("PROBE_SPLITS", "300,0,200,400").split(",")]:
    ctx.set_option(B.OPT_HOST_SPLIT_PERMILLE, split)
    for it in range(int(os.environ.get("PROBE_ITERS", "2"))):
        t = time.perf_counter()
        B.check(lib.btbb_b200_find_ac_host(ctx.h, hp.data_ptr(), n, B.LAP_ANY, 2, hits.ctypes.data, cap, C.byref(got)))
        dt = time.perf_counter() - t
        print("split", split, "iter", it, f"{dt*1e3:.1f} ms  {n/dt/1e9:.1f} Gbit/s  hits {got.value} {'ok' if got.value == want else 'MISMATCH'}", flush=True)
