"""Small, fixed workloads for ncu captures of each kernel (developer tool):

    ncu --set full -k regex:<kernel> -c 1 ... python tools/ncu_targets.py <target>

targets: v7 (promiscuous k=2), k3 / k4 / k5 (larger error tables), known (known-LAP scan),
decode0 (btbb_decode with the true clock), decode1 (64-clock sweep, full records), tc16 (compact
64-clock sweep), sieve (UAP sieve: compact sweep + candidate elimination), hops (2^27-entry hop
sequence + a winnow call), capture (pcap records of 158 005 packets formatted on the device)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from libbtbb_b200 import binding as B

target = sys.argv[1]
lib = B.lib()
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("NCU_REPS", "2"))
if target == "hops":
    ctx = B.Context(0, 0)
    d = torch.empty(1 << 27, dtype=torch.uint8, device="cuda")
    cfg = B.hop_cfg(0xA96EF25)
    for _ in range(reps):
        B.check(lib.btbb_b200_hop_sequence_dev(ctx.h, C.byref(cfg), 0, 1 << 27, d.data_ptr(), st))
    torch.cuda.synchronize()
    c, a = B.hop_winnow(ctx, cfg, 5, np.arange(12) * 7, np.arange(12) % 79)
    print("hops done", len(c))
elif target == "capture":
    n = 158005
    rng = np.random.default_rng(5)
    hits = np.zeros(n, dtype=B.HIT_DTYPE)
    hits["lap"] = rng.integers(0, 1 << 24, n)
    rec = np.zeros(n, dtype=B.DECODED_DTYPE)
    rec["payload_length"] = rng.choice([0, 17, 20, 27, 121, 183], n)
    rec["payload"] = rng.integers(0, 256, (n, 344), dtype=np.uint8)
    meta = np.zeros(n, dtype=B.PCAP_META_DTYPE)
    ctx = B.Context(0, 0)
    dh, dr, dm = (torch.from_numpy(x.view(np.uint8).copy()).cuda() for x in (hits, rec, meta))
    out = torch.empty(n * 440, dtype=torch.uint8, device="cuda")
    got = C.c_int64(0)
    for _ in range(reps):
        B.check(lib.btbb_b200_capture_records_dev(ctx.h, 0, dh.data_ptr(), dr.data_ptr(), dm.data_ptr(), n, B.LAP_ANY, 0xFF,
                                                  out.data_ptr(), out.numel(), C.byref(got), None))
    print("capture bytes", got.value)
elif target in ("v7", "k3", "k4", "k5", "known"):
    n = int(float(os.environ.get("NCU_SYMBOLS", "4e9")))
    k = {"v7": 2, "k3": 3, "k4": 4, "k5": 5, "known": 2}[target]
    cfg = B.synth_cfg(n + 72, stride=10000, mix=("ID", "DM1", "DM3", "DH1", "FHS"))
    d = torch.empty(n + 72, dtype=torch.uint8, device="cuda")
    B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), st))
    cap = n // 10000 * 4 + (1 << 22)
    hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    ctx = B.Context(0, k)
    lap = B.LAP_ANY
    if target == "known":
        c0, _ = ctx.find_ac_dev(d.data_ptr(), 1 << 24, hits.data_ptr(), cap, lap=B.LAP_ANY, k=0, stream=st)
        lap = int(hits[:1].cpu().numpy().view(B.HIT_DTYPE)["lap"][0])
    for _ in range(reps):
        cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, hits.data_ptr(), cap, lap=lap, k=k, stream=st)
    print(target, "hits", cnt, "rc", rc)
else:
    BLK, CH = 4096, 79
    blocks = int(os.environ.get("NCU_BLOCKS", "600"))
    n = blocks * CH * BLK
    coherent = target in ("tc16", "sieve")
    cfg = B.synth_cfg(n + 63, stride=BLK, ber=0.001, mix=("DM1", "DM3", "DH1", "FHS") + (("HV1",) if coherent else ()), piconets=coherent)
    d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
    B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), st))
    n_slots = blocks * CH
    truth = np.zeros((n_slots, 2), dtype=np.int32)
    pl = B.Planted()
    for sl in range(n_slots):
        lib.btbb_b200_synth_planted(C.byref(cfg), sl, C.byref(pl))
        truth[sl] = (pl.clk6, pl.uap)
    d_truth = torch.from_numpy(truth).cuda()
    cap = n_slots * 2 + 4096
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    ctx = B.Context(0, 2)
    cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=2, stream=st)
    h = d_hits[:cnt]
    off = h.view(torch.int64)[:, 0]
    if coherent:
        lap = h.view(torch.int32)[:, 2].to(torch.int64) & 0xffffff
        order = torch.sort(lap, stable=True).indices
        off, lap = off[order], lap[order]
        laps, counts = torch.unique_consecutive(lap, return_counts=True)
        gs = torch.zeros(len(laps) + 1, dtype=torch.int64, device="cuda")
        gs[1:] = torch.cumsum(counts, 0)
    pk = torch.zeros((cnt, 24), dtype=torch.uint8, device="cuda")
    pk.view(torch.int64)[:, 0] = off
    slot = off // BLK
    pk.view(torch.int32)[:, 2] = torch.clamp((slot + 1) * BLK - off, max=3125).to(torch.int32)
    pk[:, 17] = 1
    if coherent:
        pk.view(torch.int32)[:, 3] = slot.to(torch.int32)
        pk.view(torch.int32)[:, 5] = (slot % 79).to(torch.int32)
        if target == "sieve":
            states = torch.zeros((len(laps), 160), dtype=torch.uint8, device="cuda")
            for _ in range(reps):
                states.zero_()
                B.check(lib.btbb_b200_uap_sieve_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, gs.data_ptr(), len(laps),
                                                     states.data_ptr(), None, st))
        else:
            d_tc = torch.zeros(cnt * 64, dtype=torch.int16, device="cuda")
            for _ in range(reps):
                B.check(lib.btbb_b200_try_clocks_compact_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, d_tc.data_ptr(), st))
    else:
        t = d_truth[torch.clamp(slot, max=n_slots - 1)]
        pk.view(torch.int32)[:, 3] = t[:, 0]
        pk[:, 16] = t[:, 1].to(torch.uint8)
        mode = 1 if target == "decode1" else 0
        out = torch.empty((cnt * (64 if mode else 1), 372), dtype=torch.uint8, device="cuda")
        for _ in range(reps):
            B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, mode, out.data_ptr(), st))
    torch.cuda.synchronize()
    print(target, "packets", cnt)
