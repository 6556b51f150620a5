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


def test_host_packet_state_is_reused_without_changing_any_answer(orc):
    """The host path keeps the clock-independent state of the last packet and grows it on demand
    (decode_host.cpp).  Whatever the order of the questions -- short decoders first, then long ones,
    crc_check after a raw decoder, the 64-clock sweep in between -- every answer equals the one a
    fresh state gives (another packet in between evicts the entry)."""
    rng = np.random.default_rng(31337)
    other = rng.integers(0, 2, 3125, dtype=np.uint8)
    modes = [B.MODE_DECODE, B.MODE_PAYLOAD, B.MODE_CRC_CHECK] + [B.MODE_RAW + k for k in range(7)]
    for i in range(60):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        if i % 2 == 0:
            for b, d in enumerate(rng.integers(0, 1024, 183)):
                cw = orc.orc_fec23(int(d))
                sym[122 + 15 * b:122 + 15 * b + 15] = [(cw >> t) & 1 for t in range(15)]
            if i % 4 == 0:      # one uncorrectable block somewhere: fail indices must survive growing
                b = int(rng.integers(0, 60))
                sym[122 + 15 * b:122 + 15 * b + 15] ^= 1
                sym[122 + 15 * b + 3] ^= 1
        n = int(rng.choice([3125, 3125, 1500, 700, 362, 250]))
        calls = []
        for _ in range(40):
            m = int(rng.choice(modes)) | (B.MODE_FLAG_RAW_PAYLOAD if rng.integers(0, 2) else 0)
            calls.append((m, int(rng.integers(0, 64)), int(rng.integers(0, 256)), int(rng.integers(0, 16)), int(rng.integers(0, 2))))
        fresh = []
        for m, clk, uap, t, w in calls:
            B.decode_smallcall(other, 3125, 0, 0)                       # evict
            fresh.append(B.decode_smallcall(sym, n, clk, uap, whitened=w, ptype=t, mode=m)[0].tobytes())
        B.decode_smallcall(other, 3125, 0, 0)
        sweep_fresh = B.decode_smallcall(sym, n, mode=B.MODE_TRY_CLOCKS).tobytes()
        B.decode_smallcall(other, 3125, 0, 0)
        order = rng.permutation(len(calls))
        for j, k in enumerate(order):
            m, clk, uap, t, w = calls[k]
            assert B.decode_smallcall(sym, n, clk, uap, whitened=w, ptype=t, mode=m)[0].tobytes() == fresh[k], (i, k, calls[k])
            if j == 20:
                assert B.decode_smallcall(sym, n, mode=B.MODE_TRY_CLOCKS).tobytes() == sweep_fresh, i
        # a rewritten packet at the same address is a different packet
        sym2 = sym.copy()
        sym2[130:400] ^= rng.integers(0, 2, 270, dtype=np.uint8)
        m, clk, uap, t, w = calls[0]
        a = B.decode_smallcall(sym2, n, clk, uap, whitened=w, ptype=t, mode=m)[0].tobytes()
        sym[:] = sym2
        assert B.decode_smallcall(sym, n, clk, uap, whitened=w, ptype=t, mode=m)[0].tobytes() == a


def test_single_decoders_with_a_forced_type(orc):
    """fhs() / DM() / DH() / EV3() / EV4() / EV5() / HV(), btbb_decode_payload and crc_check called
    directly on a packet whose type field was set by the caller (bluetooth_packet.h:115-144) -- also a
    type that is not the decoder's own: oracle against the unmodified reference where it was built,
    host path against the oracle."""
    rng = np.random.default_rng(60606)
    R = util.ref() if util.have_ref() else None
    if R is not None:
        R.btbb_init(2)
    seen = set()
    for i, (sym, n, clk, uap, t, fn, w) in enumerate(util.forced_type_cases(orc, rng, 1500)):
        for raw in (0, 1):
            want = util.typed_one(orc, "orc", sym, n, clk, uap, t, fn, w, raw)
            if R is not None:
                _cmp(want, util.typed_one(R, "ref", sym, n, clk, uap, t, fn, w, raw), (i, "oracle vs reference", fn, t, n, raw))
            got = B.decode_smallcall(sym, n, clk, uap, whitened=w, ptype=t,
                                     mode=util.mode_of_fn(fn) | (B.MODE_FLAG_RAW_PAYLOAD if raw else 0))[0]
            _cmp(got, want, (i, "host vs oracle", fn, t, n, raw))
        seen.add((fn, int(want["rv"])))
    assert len(seen) >= 18, sorted(seen)
