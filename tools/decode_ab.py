"""A/B of decode-kernel launch configurations (developer tool): TRY_CLOCKS kernel time on the
config-3 capture for the default (24 warps x 16 staged records) and `wide` (12 x 32)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from libbtbb_b200 import binding as B
lib = B.lib()
BLK, CH, blocks = 4096, 79, 2000
n = blocks * CH * BLK
cfg = B.synth_cfg(n + 63, stride=BLK, ber=0.001, mix=("DM1", "DM3", "DH1", "FHS"))
d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()
cap = blocks * CH * 2 + 4096
d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
ctx = B.Context(0, 2)
st = torch.cuda.current_stream().cuda_stream
cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=2, stream=st)
off = d_hits[:cnt].view(torch.int64)[:, 0]
pk = torch.zeros((cnt, 24), dtype=torch.uint8, device="cuda")
pk.view(torch.int64)[:, 0] = off
pk.view(torch.int32)[:, 2] = torch.clamp((off // BLK + 1) * BLK - off, max=3125).to(torch.int32)
pk[:, 17] = 1
out = torch.empty((cnt * 64, 372), dtype=torch.uint8, device="cuda")
res, ref = {}, None
for name in sys.argv[1:] or ["default"]:
    ctx.set_option(B.OPT_DECODE_WIDE_STAGING, 1 if name == "wide" else 0)      # names: default, wide
    out.zero_()
    for _ in range(3):
        B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, 1, out.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, 1, out.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    chk = int(out.view(torch.int32).sum(dtype=torch.int64).item())
    ref = chk if ref is None else ref
    res[name] = {"ms": round(ms, 4), "Mpkt_s": round(cnt / ms / 1e3, 1), "same_output": chk == ref}
print(json.dumps(res))
