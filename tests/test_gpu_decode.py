"""Parity of the CUDA per-packet chain (unfec13/unwhiten/HEC, unfec23/CRC, all type
decoders) with the oracle and the golden fixtures, through btbb_b200_decode_host."""
import json
import os
import sys

import numpy as np
import pytest

import util
from util import B

sys.path.insert(0, util.GOLDEN)
import make_golden  # noqa: E402
from test_oracle_vs_golden import noise_type_records  # noqa: E402

pytestmark = pytest.mark.gpu


def pkts_for(planted, s, length=None, clk_delta=0, uap_delta=0, whitened=1):
    a = np.zeros(len(planted), dtype=B.PKTIN_DTYPE)
    for i, p in enumerate(planted):
        a[i]["offset"] = p.offset
        a[i]["length"] = min(3125 if length is None else length, len(s) - p.offset)
        a[i]["clkn"] = (p.clk6 + clk_delta) & 63
        a[i]["uap"] = (p.uap + uap_delta) & 255
        a[i]["whitened"] = whitened
    return a


def test_golden_decode_fixture(gpu_ctx2):
    g = json.load(open(os.path.join(util.GOLDEN, "decode.json")))
    for gs in g["streams"]:
        cfg, s = make_golden.synth_stream(gs["ber"])
        pl = util.planted_list(cfg)
        recs = gpu_ctx2.decode_host(s, pkts_for(pl, s), mode=0)
        assert recs[:4].tobytes().hex() == gs["decode_head"]
        assert util.digest(recs) == gs["decode_sha256"]
        tc = gpu_ctx2.decode_host(s, pkts_for(pl[:40], s), mode=1)
        assert util.digest(tc) == gs["try_clock_sha256"]
        odd = []
        for p in pl[:60]:
            for L in (100, 121, 122, 130, 137, 200, 361, 362, 500):
                odd.append(pkts_for([p], s, length=L)[0])
            odd.append(pkts_for([p], s, clk_delta=1)[0])
            odd.append(pkts_for([p], s, uap_delta=1)[0])
            odd.append(pkts_for([p], s, whitened=0)[0])
        got = gpu_ctx2.decode_host(s, np.array(odd, dtype=B.PKTIN_DTYPE), mode=0)
        assert util.digest(got) == gs["odd_sha256"]


def test_oracle_decode_and_try_clocks(gpu_ctx2, orc):
    cfg = B.synth_cfg(2_000_000, stride=3400, ber=0.006, seed=31337, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)
    got = gpu_ctx2.decode_host(s, pkts_for(pl, s), mode=0)
    want = np.array([util.decode_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), p.clk6, p.uap) for p in pl])
    bad = [i for i in range(len(pl)) if got[i].tobytes() != want[i].tobytes()]
    assert bad == [], (bad[:5], got[bad[0]], want[bad[0]])
    sub = pl[:120]
    got = gpu_ctx2.decode_host(s, pkts_for(sub, s), mode=1)
    want = np.array([util.try_clock_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), c)
                     for p in sub for c in range(64)])
    bad = [i for i in range(len(want)) if got[i].tobytes() != want[i].tobytes()]
    assert bad == [], (bad[:5], got[bad[0]], want[bad[0]])


def test_all_packet_types_on_noise(gpu_ctx2):
    """EV3/EV4/EV5/HV2/HV3/DV/AUX1/NULL/POLL included (fixture from the reference)."""
    g = json.load(open(os.path.join(util.GOLDEN, "noise_types.json")))
    rng = np.random.default_rng(g["seed"])
    streams, pk1, pk0, order = [], [], [], []
    for i in range(300):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        n = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        clk = int(rng.integers(0, 64))
        streams.append(sym)
        a = np.zeros(1, dtype=B.PKTIN_DTYPE)[0]
        a["offset"], a["length"], a["clkn"], a["uap"], a["whitened"] = i * 3125, n, clk, 0, 1
        pk0.append(a.copy())
    s = np.concatenate(streams)
    pk = np.array(pk0, dtype=B.PKTIN_DTYPE)
    m1 = gpu_ctx2.decode_host(s, pk, mode=1).reshape(300, 64)
    m0 = gpu_ctx2.decode_host(s, pk, mode=0)
    recs = []
    for i in range(300):
        recs.extend(m1[i, c] for c in range(0, 64, 7))
        recs.append(m0[i])
    recs = np.array(recs)
    assert len(recs) == g["count"] and util.digest(recs) == g["sha256"]


def test_header_present(gpu_ctx2, orc, product_lib):
    import ctypes as C
    import torch
    cfg = B.synth_cfg(600_000, stride=3400, ber=0.03, seed=5, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)
    pk = pkts_for(pl, s)
    pk["length"][::5] = 100
    d_s = torch.from_numpy(s).cuda()
    d_p = torch.from_numpy(pk.view(np.uint8)).cuda()
    d_r = torch.zeros(len(pl), dtype=torch.uint8, device="cuda")
    B.check(product_lib.btbb_b200_header_present_dev(gpu_ctx2.h, d_s.data_ptr(), len(s), d_p.data_ptr(),
                                                     len(pl), d_r.data_ptr(), 0))
    torch.cuda.synchronize()
    want = [orc.orc_header_present(s[p.offset:].ctypes.data, int(pk["length"][i])) for i, p in enumerate(pl)]
    assert d_r.cpu().numpy().tolist() == want and 0 < sum(want) < len(want)


def test_crafted_crc_success_paths_on_gpu(gpu_ctx2, orc):
    """EV4 / FHS (own and other clock) / DV / EV3 / EV5 CRC closures, the raw-payload flag and an
    odd output alignment, all through the kernels; every record against the oracle."""
    rng = np.random.default_rng(2024)
    cases = list(util.crafted_packets(orc, rng, 260))
    cases += [("ev35", sym, 3125, clk, uap) for sym, clk, uap in util.ev35_hunt(orc, rng, 1500)]
    s = np.concatenate([c[1] for c in cases])
    pk = np.zeros(len(cases), dtype=B.PKTIN_DTYPE)
    for i, (_, _, n, clk, uap) in enumerate(cases):
        pk[i]["offset"], pk[i]["length"], pk[i]["clkn"], pk[i]["uap"], pk[i]["whitened"] = i * 3125, n, clk, uap, 1
    got = gpu_ctx2.decode_host(s, pk, mode=0)
    raw = gpu_ctx2.decode_host(s, pk, mode=B.MODE_FLAG_RAW_PAYLOAD)
    closures = 0
    for i, (_, sym, n, clk, uap) in enumerate(cases):
        want = util.decode_one(orc, "orc", sym, 0, n, clk, uap)
        assert got[i].tobytes() == want.tobytes(), (i, got[i], want)
        assert raw[i].tobytes() == util.decode_one_raw(orc, "orc", sym, 0, n, clk, uap).tobytes(), (i, "raw")
        closures += int(want["rv"] >= 10)
    assert closures > 150
    sub = pk[:128]
    tc = gpu_ctx2.decode_host(s, sub, mode=1).reshape(len(sub), 64)
    tcr = gpu_ctx2.decode_host(s, sub, mode=1 | B.MODE_FLAG_RAW_PAYLOAD).reshape(len(sub), 64)
    for i in range(len(sub)):
        for c in range(64):
            want = util.try_clock_one(orc, "orc", cases[i][1], 0, cases[i][2], c)
            assert tc[i, c].tobytes() == want.tobytes(), (i, c, tc[i, c], want)
        # raw flag: same fields, payload bytes also where rv < 2
        assert (tcr[i]["rv"] == tc[i]["rv"]).all() and (tcr[i]["payload_length"] == tc[i]["payload_length"]).all()
        ok = tc[i]["rv"] >= 2
        assert (tcr[i]["payload"][ok] == tc[i]["payload"][ok]).all()


def test_device_entry_points_match_smallcall_and_compact(gpu_ctx2, product_lib):
    """btbb_b200_decode_dev with a misaligned stream pointer and a 4-byte (not 16-byte) aligned output
    (the bulk-store path needs 16), the compact try-clocks table, and the host small-call path: all
    the same records."""
    import ctypes as C
    import torch
    cfg = B.synth_cfg(900_000, stride=3300, ber=0.004, seed=77, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)[:200]
    pk = pkts_for(pl, s)
    ref1 = gpu_ctx2.decode_host(s, pk, mode=1).reshape(len(pl), 64)
    ref0 = gpu_ctx2.decode_host(s, pk, mode=0)
    for i in range(0, len(pl), 9):
        sym = np.ascontiguousarray(s[pl[i].offset:pl[i].offset + int(pk[i]["length"])])
        assert B.decode_smallcall(sym, len(sym), int(pk[i]["clkn"]), int(pk[i]["uap"]))[0].tobytes() == ref0[i].tobytes()
        assert B.decode_smallcall(sym, len(sym), mode=1).tobytes() == ref1[i].tobytes()
    d_raw = torch.zeros(len(s) + 64 + 5, dtype=torch.uint8, device="cuda")
    d_raw[5:5 + len(s)] = torch.from_numpy(s).cuda()
    d_p = torch.from_numpy(pk.view(np.uint8)).cuda()
    d_o = torch.zeros(len(pl) * 64 * 372 + 16, dtype=torch.uint8, device="cuda")
    for shift in (0, 4):
        B.check(product_lib.btbb_b200_decode_dev(gpu_ctx2.h, d_raw.data_ptr() + 5, len(s), d_p.data_ptr(), len(pl), 1,
                                                 d_o.data_ptr() + shift, 0))
        torch.cuda.synchronize()
        got = d_o[shift:shift + len(pl) * 64 * 372].cpu().numpy().view(B.DECODED_DTYPE).reshape(len(pl), 64)
        assert got.tobytes() == ref1.tobytes(), shift
    d_tc = torch.zeros(len(pl) * 64, dtype=torch.int16, device="cuda")
    B.check(product_lib.btbb_b200_try_clocks_compact_dev(gpu_ctx2.h, d_raw.data_ptr() + 5, len(s), d_p.data_ptr(), len(pl),
                                                         d_tc.data_ptr(), 0))
    torch.cuda.synchronize()
    tc = d_tc.cpu().numpy().view(np.uint16).reshape(len(pl), 64)
    cls = np.select([ref1["rv"] == 0, ref1["rv"] == 1, ref1["rv"] == 2, ref1["rv"] == 10], [0, 1, 2, 3], 4)
    assert ((tc & 0xff) == ref1["uap"]).all() and ((tc >> 8) == cls).all()


def test_single_decoders_with_a_forced_type_on_gpu(gpu_ctx2, orc):
    """BTBB_B200_MODE_PAYLOAD / _CRC_CHECK / _RAW + n with a caller-forced type field (matching the
    decoder or not), plain and with the raw-payload flag: every record against the oracle (which
    tests/test_decode_smallcall.py pins to the reference's exported fhs() / DM() / ... for these calls)."""
    rng = np.random.default_rng(60606)
    cases = list(util.forced_type_cases(orc, rng, 1800))
    s = np.concatenate([c[0] for c in cases])
    by_fn = {}
    for i, c in enumerate(cases):
        by_fn.setdefault(c[5], []).append(i)
    checked = 0
    for fn, idx in sorted(by_fn.items()):
        pk = np.zeros(len(idx), dtype=B.PKTIN_DTYPE)
        for j, i in enumerate(idx):
            _, n, clk, uap, t, _, w = cases[i]
            pk[j]["offset"], pk[j]["length"], pk[j]["clkn"], pk[j]["uap"], pk[j]["whitened"], pk[j]["type"] = i * 3125, n, clk, uap, w, t
        for raw in (0, 1):
            got = gpu_ctx2.decode_host(s, pk, mode=util.mode_of_fn(fn) | (B.MODE_FLAG_RAW_PAYLOAD if raw else 0))
            for j, i in enumerate(idx):
                sym, n, clk, uap, t, _, w = cases[i]
                want = util.typed_one(orc, "orc", sym, n, clk, uap, t, fn, w, raw)
                assert got[j].tobytes() == want.tobytes(), (fn, i, t, n, raw, got[j], want)
                checked += 1
    assert checked == 2 * len(cases) and len(by_fn) == 9
