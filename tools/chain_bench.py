"""BASELINE configs[2]: full chain on a 79-channel interleaved capture.

    python tools/chain_bench.py [--blocks N] > profiles/r01_chain.json

Capture layout [block][79 channels][4096 symbols] (SURVEY.md 8d cfg 3): every 4096-symbol
channel block carries one planted packet (DM1 / DM3 / DH1 / FHS).  Chain: find_ac (promiscuous,
k=2) -> per hit the 64-clock try_clock + crc_check sweep (the inner loop of
bluetooth_piconet.c:675-689) and the known-clock btbb_decode_header + btbb_decode_payload.
Everything stays on the device; a sample is checked against the oracle."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import util
from util import B

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=2000)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
lib = B.lib()
BLK, CH = 4096, 79
n = a.blocks * CH * BLK
cfg = B.synth_cfg(n + 63, stride=BLK, ber=0.001, mix=("DM1", "DM3", "DH1", "FHS"))
d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()
cap = a.blocks * CH * 2 + 4096
d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
ctx = B.Context(0, 2)
st = torch.cuda.current_stream().cuda_stream


# ground truth per 4096-symbol slot (clock / UAP the planted packet was whitened / HEC'ed with)
n_slots = a.blocks * CH
truth = np.zeros((n_slots, 2), dtype=np.int32)
for sl in range(n_slots):
    p = B.planted(cfg, sl)
    truth[sl] = (p.clk6, p.uap)
d_truth = torch.from_numpy(truth).cuda()
out_buf = torch.empty((cap * 64, 372), dtype=torch.uint8, device="cuda") if cap * 64 * 372 < 40e9 else None


def chain(mode, true_clock=False):
    cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=2, stream=st)
    off = d_hits[:cnt].view(torch.int64)[:, 0]
    pk = torch.zeros((cnt, 24), dtype=torch.uint8, device="cuda")
    pk.view(torch.int64)[:, 0] = off
    length = torch.clamp((off // BLK + 1) * BLK - off, max=3125).to(torch.int32)   # symbols left in the channel block
    pk.view(torch.int32)[:, 2] = length
    pk[:, 17] = 1                                                                  # whitened
    if true_clock:
        t = d_truth[torch.clamp(off // BLK, max=n_slots - 1)]
        pk.view(torch.int32)[:, 3] = t[:, 0]
        pk[:, 16] = t[:, 1].to(torch.uint8)
    out = out_buf[: cnt * (64 if mode == 1 else 1)]
    B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, mode, out.data_ptr(), st))
    return cnt, pk, out


def kernel_only(mode, pk, cnt, out, iters=10):
    """the decode kernel alone on a fixed packet list (CUDA events on its stream)"""
    for _ in range(2):
        B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, mode, out.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, mode, out.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"workload": f"{a.blocks} blocks x 79 channels x 4096 symbols, one packet per channel block, BER 0.1%",
       "symbols": n}
for mode, tc, name in ((1, False, "find_ac + 64-clock try_clock/crc_check sweep"),
                       (0, True, "find_ac + decode with the true clock / UAP (header + payload + CRC)"),
                       (0, False, "find_ac + decode (clock 0, UAP 0: header reject path)")):
    for _ in range(2):
        cnt, pk, out = chain(mode, tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        cnt, pk, out = chain(mode, tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    kms = kernel_only(mode, pk, cnt, out)
    pkh = pk.cpu().numpy().reshape(-1).view(B.PKTIN_DTYPE)
    rec = out[:cnt].cpu().numpy().reshape(-1).view(B.DECODED_DTYPE) if mode == 0 else None
    sym_bytes = int(np.minimum(pkh["length"], 3125).sum())
    out_bytes = cnt * 372 * (64 if mode == 1 else 1)
    res[name] = {"ms": ms, "packets": int(cnt), "packets_per_s": cnt / (ms / 1e3), "gbit_s": n / (ms / 1e3) / 1e9,
                 "decode_kernel_ms": kms, "decode_kernel_packets_per_s": cnt / (kms / 1e3),
                 "decode_kernel_GBps_symbols_plus_records": (sym_bytes + out_bytes) / (kms / 1e3) / 1e9,
                 "symbol_bytes": sym_bytes, "record_bytes": out_bytes}
    if rec is not None:
        res[name]["rv_histogram"] = {str(int(k)): int(v) for k, v in zip(*np.unique(rec["rv"], return_counts=True))}
# ---- UAP / CLK1-6 discovery on a piconet-coherent capture: find_ac -> group hits by LAP
# -> btbb_b200_uap_sieve_dev.  CLKN of a packet = its 4096-symbol slot index, channel = slot % 79. ----
cfg2 = B.synth_cfg(n + 63, stride=BLK, ber=0.001, mix=("DM1", "DM3", "DH1", "FHS", "HV1"), piconets=True)
B.check(lib.btbb_b200_synth_dev(C.byref(cfg2), d.data_ptr(), 0)); torch.cuda.synchronize()


def sieve_chain():
    cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=2, stream=st)
    h = d_hits[:cnt]
    off = h.view(torch.int64)[:, 0]
    lap = h.view(torch.int32)[:, 2].to(torch.int64) & 0xffffff
    order = torch.sort(lap, stable=True).indices                     # group by LAP, arrival order kept
    off, lap = off[order], lap[order]
    laps, counts = torch.unique_consecutive(lap, return_counts=True)
    gs = torch.zeros(len(laps) + 1, dtype=torch.int64, device="cuda")
    gs[1:] = torch.cumsum(counts, 0)
    pk = torch.zeros((cnt, 24), dtype=torch.uint8, device="cuda")
    pk.view(torch.int64)[:, 0] = off
    slot = off // BLK
    pk.view(torch.int32)[:, 2] = torch.clamp((slot + 1) * BLK - off, max=3125).to(torch.int32)
    pk.view(torch.int32)[:, 3] = slot.to(torch.int32)                 # clkn
    pk[:, 17] = 1
    pk.view(torch.int32)[:, 5] = (slot % 79).to(torch.int32)          # channel in `reserved`
    states = torch.zeros((len(laps), 160), dtype=torch.uint8, device="cuda")
    rv = torch.zeros(cnt, dtype=torch.int8, device="cuda")
    B.check(lib.btbb_b200_uap_sieve_dev(ctx.h, d.data_ptr(), n + 63, pk.data_ptr(), cnt, gs.data_ptr(), len(laps),
                                         states.data_ptr(), rv.data_ptr(), st))
    return cnt, pk, gs, laps, states, rv


for _ in range(2):
    sieve_chain()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    cnt, pk, gs, laps, states, rv = sieve_chain()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
stn = states.cpu().numpy().reshape(-1).view(B.SIEVE_DTYPE)
res["find_ac + UAP/CLK1-6 sieve (btbb_uap_from_header per piconet)"] = {
    "ms": ms, "packets": int(cnt), "piconets": int(len(laps)), "piconets_resolved": int(((stn["flags"] >> 2) & 1).sum()),
    "packets_per_s": cnt / (ms / 1e3), "gbit_s": n / (ms / 1e3) / 1e9}
# parity of the whole sieve result against the oracle, and the reference C path's time for the same packets
s_host = d.cpu().numpy()
pk_h = pk.cpu().numpy().reshape(-1).view(B.PKTIN_DTYPE)
gs_h = gs.cpu().numpy()
t0 = time.time()
want_st, want_rv = util.sieve_run(util.ref() if util.have_ref() else util.oracle(), "ref" if util.have_ref() else "orc", s_host, pk_h, gs_h)
t_cpu = time.time() - t0
res["sieve_parity"] = {"states_match": bool(want_st.tobytes() == stn.tobytes()), "rv_match": bool(want_rv.tobytes() == rv.cpu().numpy().tobytes()),
                       "cpu_kind": "reference" if util.have_ref() else "port", "cpu_seconds_same_packets_1_thread": round(t_cpu, 3)}
del s_host
B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()

# parity sample: first 300 hits, 64 clocks each, against the oracle
cnt, pk, out = chain(1)
torch.cuda.synchronize()
O = util.oracle()
s = d[: 400 * BLK + 4000].cpu().numpy()
pkh = pk[:300].cpu().numpy().reshape(-1).view(B.PKTIN_DTYPE)
got = out[: 300 * 64].cpu().numpy().reshape(-1).view(B.DECODED_DTYPE)
bad = 0
crc_ok = 0
for i in range(300):
    for c in range(0, 64, 9):
        w = util.try_clock_one(O, "orc", s, int(pkh[i]["offset"]), int(pkh[i]["length"]), c)
        bad += w.tobytes() != got[i * 64 + c].tobytes()
    crc_ok += int((got[i * 64:(i + 1) * 64]["rv"] >= 10).any())
res["parity_sample"] = {"records_checked": 300 * 8, "mismatches": int(bad), "packets_with_a_crc_clean_clock": crc_ok}
print(json.dumps(res))
