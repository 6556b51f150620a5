"""BASELINE configs[2] end to end on the GPU, through the C ABI: access codes of a 79-channel
interleaved capture ([block][79][4096], SURVEY.md 8d cfg 3) -> per hit btbb_decode_header +
btbb_decode_payload with the true clock / UAP -> the 64-clock try_clock + crc_check sweep ->
btbb_uap_from_header per piconet.  Every record is compared with the oracle, and the digests with
the fixture the unmodified reference produced (tests/golden/chain79.json, make_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import util
from util import B

pytestmark = pytest.mark.gpu


def test_chain_on_79_channel_capture(gpu_ctx2, orc, product_lib):
    import torch
    g = json.load(open(os.path.join(util.GOLDEN, "chain79.json")))
    cfg, s, n = util.chain79_case(g["blocks"])
    assert n == g["symbols"]
    lib, ctx = product_lib, gpu_ctx2
    # the capture is generated on the device too, and everything below stays there
    d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
    B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), s)
    cap = 4096
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, lap=B.LAP_ANY, k=2)
    assert rc == 0 and cnt == g["hits"]
    hits = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
    assert orc.orc_init(2) == 0
    assert hits.tobytes() == util.find_all(orc, "orc", s, n, B.LAP_ANY, 2).tobytes()
    assert util.digest(hits) == g["hits_sha256"]

    dec, sv, gs, laps = util.chain79_packets(cfg, hits)
    d_pk = torch.from_numpy(dec.view(np.uint8)).cuda()
    d_out = torch.zeros((cnt * 64, 372), dtype=torch.uint8, device="cuda")

    def run(mode, count=cnt):
        B.check(lib.btbb_b200_decode_dev(ctx.h, d.data_ptr(), n + 63, d_pk.data_ptr(), count, mode, d_out.data_ptr(), 0))
        torch.cuda.synchronize()
        return d_out[: count * (64 if (mode & 0xff) == 1 else 1)].cpu().numpy().reshape(-1).view(B.DECODED_DTYPE).copy()

    # --- decode with the true clock / UAP ---
    recs = run(B.MODE_DECODE)
    for i, p in enumerate(dec):
        want = util.decode_one(orc, "orc", s, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"]))
        assert recs[i].tobytes() == want.tobytes(), (i, recs[i], want)
    assert util.digest(recs) == g["decode_sha256"]
    hist = {str(k): int(v) for k, v in zip(*np.unique(recs["rv"], return_counts=True))}
    assert hist == g["rv_hist"] and hist.get("10", 0) > 150 and hist.get("1000", 0) > 50
    raw = run(B.MODE_DECODE | B.MODE_FLAG_RAW_PAYLOAD)
    assert util.digest(raw) == g["decode_raw_sha256"]
    # --- 64-clock sweep ---
    tc = run(B.MODE_TRY_CLOCKS)
    for i in range(0, cnt, 7):
        for c in range(64):
            want = util.try_clock_one(orc, "orc", s, int(dec[i]["offset"]), int(dec[i]["length"]), c)
            assert tc[i * 64 + c].tobytes() == want.tobytes(), (i, c)
    assert util.digest(tc) == g["try_clocks_sha256"]
    # --- UAP / CLK1-6 discovery per piconet ---
    d_sv = torch.from_numpy(sv.view(np.uint8)).cuda()
    d_gs = torch.from_numpy(gs).cuda()
    d_st = torch.zeros((len(gs) - 1, 160), dtype=torch.uint8, device="cuda")
    d_rv = torch.zeros(cnt, dtype=torch.int8, device="cuda")
    B.check(lib.btbb_b200_uap_sieve_dev(ctx.h, d.data_ptr(), n + 63, d_sv.data_ptr(), cnt, d_gs.data_ptr(), len(gs) - 1,
                                         d_st.data_ptr(), d_rv.data_ptr(), 0))
    torch.cuda.synchronize()
    st = d_st.cpu().numpy().reshape(-1).view(B.SIEVE_DTYPE)
    rv = d_rv.cpu().numpy()
    want_st, want_rv = util.sieve_run(orc, "orc", s, sv, gs)
    assert st.tobytes() == want_st.tobytes() and rv.tobytes() == want_rv.tobytes()
    assert [util.digest(st), util.digest(rv)] == g["sieve_sha256"]
    assert int(((st["flags"] >> 2) & 1).sum()) == g["piconets_resolved"] == g["piconets"]
    # the discovered UAP / clock are the planted ones
    for gi, lap in enumerate(laps):
        p = next(B.planted(cfg, int(q["offset"]) // util.CHAIN_BLK) for q in sv[gs[gi]:gs[gi + 1]])
        assert st[gi]["uap"] == p.uap


def test_gpu_records_to_pcap_files(gpu_ctx2, orc, product_lib, tmp_path):
    """GPU hits + GPU decode records (raw-payload flag) -> btbb_b200_pcap_bredr_records /
    btbb_b200_pcapng_bredr_blocks == the files the reference writes for the same capture, EVERY packet
    included; digests from the reference in tests/golden/chain79.json."""
    import hashlib
    import torch
    import test_pcap
    g = json.load(open(os.path.join(util.GOLDEN, "chain79.json")))
    cfg, s, n = util.chain79_case(g["blocks"])
    d = torch.from_numpy(s).cuda()
    cap = 4096
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, lap=B.LAP_ANY, k=2)
    hits = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
    dec, sv, gs, laps = util.chain79_packets(cfg, hits)
    d_pk = torch.from_numpy(dec.view(np.uint8)).cuda()
    d_out = torch.zeros((cnt, 372), dtype=torch.uint8, device="cuda")
    B.check(product_lib.btbb_b200_decode_dev(gpu_ctx2.h, d.data_ptr(), n + 63, d_pk.data_ptr(), cnt,
                                             B.MODE_DECODE | B.MODE_FLAG_RAW_PAYLOAD, d_out.data_ptr(), 0))
    torch.cuda.synchronize()
    recs = d_out.cpu().numpy().reshape(-1).view(B.DECODED_DTYPE)
    meta = util.chain79_meta(dec)
    pcap = B.pcap_bredr(hits, recs, meta)
    png = B.pcapng_bredr_blocks(hits, recs, meta)
    assert [len(pcap), hashlib.sha256(pcap).hexdigest()] == g["pcap"]
    assert [len(png), hashlib.sha256(png).hexdigest()] == g["pcapng_blocks"]
    if util.have_ref():
        assert pcap == test_pcap._ref_pcap(tmp_path, s, hits, dec, meta)


def test_capture_records_formatted_on_the_device(gpu_ctx2, product_lib):
    """btbb_b200_capture_records_dev == the host formatters, byte for byte: GPU hit and decode records of
    the 79-channel capture (failed decodes included), then records with every payload length 0..400+ and
    odd metadata so that each padding / flag case of both formats occurs; size query, short buffer, n = 0."""
    import ctypes as C
    import torch
    g = json.load(open(os.path.join(util.GOLDEN, "chain79.json")))
    cfg, s, n = util.chain79_case(g["blocks"])
    d = torch.from_numpy(s).cuda()
    cap = 4096
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, lap=B.LAP_ANY, k=2)
    hits = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
    dec, sv, gs, laps = util.chain79_packets(cfg, hits)
    d_pk = torch.from_numpy(dec.view(np.uint8)).cuda()
    d_rec = torch.zeros((cnt, 372), dtype=torch.uint8, device="cuda")
    B.check(product_lib.btbb_b200_decode_dev(gpu_ctx2.h, d.data_ptr(), n + 63, d_pk.data_ptr(), cnt,
                                             B.MODE_DECODE | B.MODE_FLAG_RAW_PAYLOAD, d_rec.data_ptr(), 0))
    torch.cuda.synchronize()
    recs = d_rec.cpu().numpy().reshape(-1).view(B.DECODED_DTYPE)
    meta = util.chain79_meta(dec)

    def on_device(hits_h, recs_h, meta_h, fmt, reflap=B.LAP_ANY, refuap=0xFF, short=False):
        dh = torch.from_numpy(np.ascontiguousarray(hits_h).view(np.uint8).copy()).cuda()
        dr = torch.from_numpy(np.ascontiguousarray(recs_h).view(np.uint8).copy()).cuda()
        dm = torch.from_numpy(np.ascontiguousarray(meta_h).view(np.uint8).copy()).cuda()
        need = C.c_int64(-1)
        B.check(product_lib.btbb_b200_capture_records_dev(gpu_ctx2.h, fmt, dh.data_ptr(), dr.data_ptr(), dm.data_ptr(), len(hits_h),
                                                          reflap, refuap, None, 0, C.byref(need), None))
        out = torch.full((max(need.value, 1) + 64,), 0xAB, dtype=torch.uint8, device="cuda")
        got = C.c_int64(-1)
        give = need.value - 1 if short else need.value
        B.check(product_lib.btbb_b200_capture_records_dev(gpu_ctx2.h, fmt, dh.data_ptr(), dr.data_ptr(), dm.data_ptr(), len(hits_h),
                                                          reflap, refuap, out.data_ptr(), give, C.byref(got), None))
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        assert got.value == need.value
        if short:
            assert (o == 0xAB).all()                       # nothing written when it does not fit
            return None
        assert (o[need.value:] == 0xAB).all()              # nothing written past the end
        return o[:need.value].tobytes()

    assert on_device(hits, recs, meta, 0) == B.pcap_bredr(hits, recs, meta)[24:]      # pcap_bredr = file header + records
    assert on_device(hits, recs, meta, 1) == B.pcapng_bredr_blocks(hits, recs, meta)
    # synthetic records: every payload length, both sides of every branch of the flag word
    rng = np.random.default_rng(12)
    m = 3000
    h2 = np.zeros(m, dtype=B.HIT_DTYPE)
    h2["offset"] = np.arange(m) * 4000
    h2["lap"] = rng.integers(0, 1 << 24, m)
    h2["ac_errors"] = rng.integers(0, 6, m)
    r2 = np.zeros(m, dtype=B.DECODED_DTYPE)
    r2["payload_length"] = np.concatenate([np.arange(0, 420), rng.integers(-3, 420, m - 420)])
    r2["payload"] = rng.integers(0, 256, (m, 344), dtype=np.uint8)
    r2["header_packed"] = rng.integers(0, 1 << 18, m)
    m2 = np.zeros(m, dtype=B.PCAP_META_DTYPE)
    m2["ns"] = rng.integers(0, 1 << 62, m, dtype=np.uint64)
    m2["sigdbm"] = rng.integers(-100, 10, m)
    m2["noisedbm"] = rng.integers(-100, 10, m)
    m2["channel"] = rng.integers(0, 79, m)
    m2["transport"] = rng.integers(0, 4, m)
    m2["modulation"] = rng.integers(0, 3, m)
    for reflap, refuap in ((B.LAP_ANY, 0xFF), (0x9E8B33, 0x47), (0x123456, 0xFF)):
        assert on_device(h2, r2, m2, 0, reflap, refuap) == B.pcap_bredr(h2, r2, m2, reflap, refuap)[24:]
        assert on_device(h2, r2, m2, 1, reflap, refuap) == B.pcapng_bredr_blocks(h2, r2, m2, reflap, refuap)
    for k in (1, 7, 1023, 1024, 1025, 2049):
        assert on_device(h2[:k], r2[:k], m2[:k], 1) == B.pcapng_bredr_blocks(h2[:k], r2[:k], m2[:k])
    assert on_device(h2, r2, m2, 0, short=True) is None
    z = C.c_int64(-1)
    B.check(product_lib.btbb_b200_capture_records_dev(gpu_ctx2.h, 0, None, None, None, 0, B.LAP_ANY, 0xFF, None, 0, C.byref(z), None))
    assert z.value == 0
