#!/bin/bash
# round 2, job e (1 GPU): full -m gpu suite (hops, pcap e2e, decode variants) + decode A/B + hop kernel timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
python tools/decode_ab.py default full12 2>&1 | tail -1
python - <<'PY'
import ctypes as C, sys, time
sys.path.insert(0, '.')
import torch
from libbtbb_b200 import binding as B
lib = B.lib(); ctx = B.Context(0, 2)
d = torch.empty(1 << 27, dtype=torch.uint8, device='cuda')
cfg = B.hop_cfg(0xA96EF25)
for _ in range(2): B.check(lib.btbb_b200_hop_sequence_dev(ctx.h, C.byref(cfg), 0, 1 << 27, d.data_ptr(), 0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): B.check(lib.btbb_b200_hop_sequence_dev(ctx.h, C.byref(cfg), 0, 1 << 27, d.data_ptr(), 0))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("hop_sequence 2^27 entries: %.3f ms = %.0f GB/s written" % (ms, (1 << 27) / ms / 1e6))
import numpy as np
t0 = time.perf_counter()
for _ in range(20): c, a = B.hop_winnow(ctx, cfg, 5, np.arange(12) * 7, np.arange(12) % 79)
print("hop_winnow (12 observations, host arrays in/out): %.3f ms per call" % ((time.perf_counter() - t0) / 20 * 1e3))
PY
