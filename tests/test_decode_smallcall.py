"""The per-packet chain's arithmetic (libbtbb_b200/csrc/decode_core.h) checked on the CPU.

decode_core.h is compiled twice: into the kernels of decode.cu and, for the classic single-packet
calls, into btbb_b200_decode_smallcall (host).  The host build runs here against the oracle on
planted packets, truncated / wrong-clock / unwhitened variants and on noise that reaches every
packet-type decoder, so the table-driven CRC algebra, the EV3/EV4/EV5 length searches and the
record layout are pinned without a GPU; tests/test_gpu_decode.py pins the kernels themselves."""
import json
import os
import sys

import numpy as np
import pytest

import util
from util import B

sys.path.insert(0, util.GOLDEN)
import make_golden  # noqa: E402


def _cmp(got, want, what):
    assert got.tobytes() == want.tobytes(), (what, got, want)


def test_planted_packets_all_modes(orc):
    cfg = B.synth_cfg(1_200_000, stride=3400, ber=0.004, seed=4242, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)
    assert len(pl) > 300
    rng = np.random.default_rng(7)
    for i, p in enumerate(pl):
        L = min(3125, len(s) - p.offset)
        sym = np.ascontiguousarray(s[p.offset:p.offset + L])
        _cmp(B.decode_smallcall(sym, L, p.clk6, p.uap)[0], util.decode_one(orc, "orc", s, p.offset, L, p.clk6, p.uap), (i, "decode"))
        _cmp(B.decode_smallcall(sym, L, p.clk6, p.uap, mode=B.MODE_FLAG_RAW_PAYLOAD)[0],
             util.decode_one_raw(orc, "orc", s, p.offset, L, p.clk6, p.uap), (i, "decode raw"))
        if i % 4 == 0:
            L2 = int(rng.choice([100, 121, 122, 130, 137, 200, 361, 362, 500, 1000, 2000]))
            L2 = min(L2, L)
            _cmp(B.decode_smallcall(sym, L2, p.clk6, p.uap)[0], util.decode_one(orc, "orc", s, p.offset, L2, p.clk6, p.uap), (i, "short", L2))
            _cmp(B.decode_smallcall(sym, L, p.clk6 + 1, p.uap)[0], util.decode_one(orc, "orc", s, p.offset, L, p.clk6 + 1, p.uap), (i, "clk+1"))
            _cmp(B.decode_smallcall(sym, L, p.clk6, p.uap, whitened=0)[0],
                 util.decode_one(orc, "orc", s, p.offset, L, p.clk6, p.uap, whitened=0), (i, "unwhitened"))
        if i % 16 == 0:
            got = B.decode_smallcall(sym, L, mode=B.MODE_TRY_CLOCKS)
            for c in range(64):
                _cmp(got[c], util.try_clock_one(orc, "orc", s, p.offset, L, c), (i, "try_clock", c))


def test_all_packet_types_on_noise_match_reference_fixture():
    """Same records as tests/test_gpu_decode.py::test_all_packet_types_on_noise (fixture from the reference)."""
    g = json.load(open(os.path.join(util.GOLDEN, "noise_types.json")))
    rng = np.random.default_rng(g["seed"])
    recs = []
    for i in range(300):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        n = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        clk = int(rng.integers(0, 64))
        m1 = B.decode_smallcall(sym, n, mode=B.MODE_TRY_CLOCKS)
        recs.extend(m1[c] for c in range(0, 64, 7))
        recs.append(B.decode_smallcall(sym, n, clk, 0)[0])
    recs = np.array(recs)
    assert len(recs) == g["count"] and util.digest(recs) == g["sha256"]


def test_noise_raw_payload_and_single_decoders(orc):
    """Failed decodes leave the reference's pkt->payload bytes (raw flag); every type through
    btbb_decode_payload with a forced type / UAP."""
    rng = np.random.default_rng(99)
    for i in range(400):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        if i % 3 == 0:      # valid FEC 2/3 blocks so that DM / EV4 / FHS get past unfec23
            for b, d in enumerate(rng.integers(0, 1024, 183)):
                cw = orc.orc_fec23(int(d))
                sym[122 + 15 * b:122 + 15 * b + 15] = [(cw >> t) & 1 for t in range(15)]
        n = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        clk = int(rng.integers(0, 64))
        # find the UAP that makes btbb_decode_header accept this clock, so every type's payload path runs
        t = util.try_clock_one(orc, "orc", sym, 0, n, clk)
        uap = int(t["uap"])
        _cmp(B.decode_smallcall(sym, n, clk, uap, mode=B.MODE_FLAG_RAW_PAYLOAD)[0],
             util.decode_one_raw(orc, "orc", sym, 0, n, clk, uap), (i, "raw"))
        _cmp(B.decode_smallcall(sym, n, clk, uap)[0], util.decode_one(orc, "orc", sym, 0, n, clk, uap), (i, "plain"))


from util import crafted_packets  # noqa: E402


def test_crafted_crc_success_paths(orc):
    rng = np.random.default_rng(2024)
    seen = {"ev4": 0, "fhs_own": 0, "fhs_other": 0, "dv": 0, "ev35": 0, "ev4_fecfail": 0}
    for i, (kind, sym, n, clk, uap) in enumerate(crafted_packets(orc, rng, 260)):
        want = util.decode_one(orc, "orc", sym, 0, n, clk, uap)
        _cmp(B.decode_smallcall(sym, n, clk, uap)[0], want, (i, "crafted"))
        _cmp(B.decode_smallcall(sym, n, clk, uap, mode=B.MODE_FLAG_RAW_PAYLOAD)[0],
             util.decode_one_raw(orc, "orc", sym, 0, n, clk, uap), (i, "crafted raw"))
        tc = B.decode_smallcall(sym, n, mode=B.MODE_TRY_CLOCKS)
        for c in (clk, (clk + 17) & 63, (clk + 40) & 63):
            _cmp(tc[c], util.try_clock_one(orc, "orc", sym, 0, n, c), (i, "crafted try_clock", c))
        assert want["header_ok"] == 1
        if want["rv"] >= 10:
            seen[kind] += 1
        if kind == "ev4" and want["rv"] in (0, 1):
            seen["ev4_fecfail"] += 1
    # EV3 / EV5 CRC closures are luck: hunt for them with the oracle's own verdicts
    for i, (sym, clk, uap) in enumerate(util.ev35_hunt(orc, rng, 3000)):
        want = util.decode_one(orc, "orc", sym, 0, 3125, clk, uap)
        _cmp(B.decode_smallcall(sym, 3125, clk, uap)[0], want, (i, "ev35"))
        seen["ev35"] += int(want["rv"] == 10)
    assert seen["ev4"] > 25 and seen["fhs_own"] > 20 and seen["fhs_other"] > 20 and seen["dv"] > 40, seen
    assert seen["ev35"] >= 3 and seen["ev4_fecfail"] > 5, seen
