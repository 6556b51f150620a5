"""Pins the CPU oracle (oracle/oracle.c) to the reference:
  (a) the known-answer vectors of the reference's own tests/ directory
      (tests/test_syndromes.c:38-74, tests/test_fec23.c:38-86, tests/test_header.c:22-45),
  (b) the fixtures tests/golden/*.json, whose outputs were produced by the unmodified
      reference (tests/golden/make_golden.py),
  (c) when oracle/_ref/libbtbb_ref.so is present, a live differential run.
CPU only."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from util import B

sys.path.insert(0, util.GOLDEN)
import make_golden  # noqa: E402  (stream builders shared with the fixture generator)


def load(name):
    return json.load(open(os.path.join(util.GOLDEN, name)))


def cbits(s):
    return np.array([int(c) for c in s], dtype=np.uint8)


# ---- (a) the reference's own known-answer vectors ----
def test_reference_kat_syndromes(orc):
    assert orc.orc_syndrome(0xcc7b7268ff614e1b) == 0                 # test_syndromes.c:38-47
    assert orc.orc_syndrome(0xcc7d7268ff614e1b) == 0x299c6f9b5       # test_syndromes.c:41,50


def test_reference_kat_syncword_correction(orc):
    # test_syndromes.c:62-74: DEFAULT_AC and a one-bit-error copy both decode to the LAP 0xffffff word
    pn = 0x83848D96BBCC54FC
    assert orc.orc_gen_syncword(0xffffff) == 0x4ffffffe44ad1ae7 == 0xcc7b7268ff614e1b ^ pn
    assert orc.orc_init(2) == 0
    for cw in (0xcc7b7268ff614e1b, 0xcc7b7268ff514e1b):
        sw = cw ^ pn
        stream = np.array([(sw >> i) & 1 for i in range(64)] + [0] * 64, dtype=np.uint8)
        h = util.find_all(orc, "orc", stream, 1, B.LAP_ANY, 2)
        assert len(h) == 1 and h[0]["lap"] == 0xffffff and h[0]["offset"] == 0


def test_reference_kat_syncwords(orc):
    for lap, sw in ((0, 0xb0000002c7820e7e), (0xffffff, 0x4ffffffe44ad1ae7), (0x9e8b33, 0x4e7a2cce331a3ae2)):
        assert orc.orc_gen_syncword(lap) == sw


FEC23_PARITY = ["11010", "01101", "11100", "01110", "00111", "11001", "10110", "01011", "11111", "10101"]


def test_reference_kat_fec23(orc):
    # test_fec23.c:38-86: 10 clean blocks and the same blocks with the data bit knocked out
    for i, par in enumerate(FEC23_PARITY):
        data = [0] * 10
        data[i] = 1
        for rx in (data, [0] * 10):
            blk = np.array(rx + [int(c) for c in par], dtype=np.uint8)
            out = np.zeros(10, dtype=np.uint8)
            assert orc.orc_unfec23(blk.ctypes.data, 1, out.ctypes.data) == 1
            assert out.tolist() == data


HEADER_KATS = """00 123 e1 770007 007070 000777
47 123 06 770007 007007 700000
00 124 32 007007 007007 007700
47 124 d5 007007 007070 707077
00 125 5a 707007 007007 077070
47 125 bd 707007 007070 777707
00 126 e2 077007 007007 000777
47 126 05 077007 007070 700000
00 127 8a 777007 007007 070007
47 127 6d 777007 007070 770770
00 11b 9e 770770 007007 777007
47 11b 79 770770 007070 077770
00 11c 4d 007770 007070 770070
47 11c aa 007770 007007 070707
00 11d 25 707770 007070 700700
47 11d c2 707770 007007 000077
00 11e 9d 077770 007070 777007
47 11e 7a 077770 007007 077770
00 11f f5 777770 007070 707777
47 11f 12 777770 007007 007000"""


def test_reference_kat_hec(orc):
    # test_header.c:22-45: UAP, 10 header bits, HEC and the FEC-1/3 expansion in octal
    for line in HEADER_KATS.splitlines():
        uap, data, hec, *octal = line.split()
        uap, data, hec = int(uap, 16), int(data, 16), int(hec, 16)
        assert orc.orc_hec(data, uap) == hec
        assert orc.orc_uap_from_hec(data, hec) == uap
        sym = []
        for digit in "".join(octal):
            sym += [(int(digit) >> 2) & 1, (int(digit) >> 1) & 1, int(digit) & 1]
        sym = np.array(sym, dtype=np.uint8)
        out = np.zeros(18, dtype=np.uint8)
        assert orc.orc_unfec13(sym.ctypes.data, out.ctypes.data, 18) == 1
        word = sum(int(b) << i for i, b in enumerate(out))
        assert word == (data | (hec << 10))


# ---- (b) golden fixtures ----
def test_golden_primitives(orc):
    g = load("primitives.json")
    for cw, syn in g["syndrome"]:
        assert orc.orc_syndrome(int(cw, 16)) == int(syn, 16)
    for lap, sw in g["syncword"]:
        assert orc.orc_gen_syncword(lap) == int(sw, 16)
    assert [orc.orc_barker_distance(b) for b in range(128)] == g["barker_distance"]
    for b in range(128):
        if g["barker_distance"][b] <= 1:   # the reference only consults barker_correct after the <=1 filter
            assert orc.orc_barker_correct(b) == int(g["barker_correct"][b], 16)
    for clk in range(64):
        assert "".join(str(orc.orc_whiten_bit(clk, i)) for i in range(127)) == g["whitening"][clk]
    assert [orc.orc_fec23(d) for d in range(1024)] == g["fec23"]
    for d, h, u in g["uap_from_hec"]:
        assert orc.orc_uap_from_hec(d, h) == u
    for bits, uap, crc in g["crc"]:
        a = cbits(bits) if bits else np.zeros(1, dtype=np.uint8)
        assert orc.orc_crc16(a.ctypes.data, len(bits), uap) == crc
    for bits, L, ok, out in g["unfec13"]:
        a, o = cbits(bits), np.zeros(L, dtype=np.uint8)
        assert orc.orc_unfec13(a.ctypes.data, o.ctypes.data, L) == ok
        assert "".join(map(str, o)) == out
    for bits, L, ok, out in g["unfec23"]:
        a, o = cbits(bits), np.zeros(((L + 9) // 10) * 10, dtype=np.uint8)
        assert orc.orc_unfec23(a.ctypes.data, L, o.ctypes.data) == ok
        if ok:
            assert "".join(map(str, o)) == out
    assert g["sizeof_packet"] == 5952


_FIND_SNIPPET = """
import sys, json
sys.path.insert(0, {tests!r}); sys.path.insert(0, {golden!r})
import util, make_golden
from util import B
k_init = int(sys.argv[1])
O = util.oracle(); assert O.orc_init(k_init) == 0
cases = json.load(open({golden!r} + '/find_ac.json'))[k_init]['cases']
s1 = make_golden.find_ac_stream(1234, 1 << 20); cfg, s2 = make_golden.synth_stream(0.005)
bad = []
for c in cases:
    s = s1 if c['stream'] == 'rand1234' else s2
    h = util.find_all(O, 'orc', s, c['n'], c['lap'], c['k'])
    if len(h) != c['count'] or util.digest(h) != c['sha256']:
        bad.append((c['stream'], c['lap'], c['k'], len(h), c['count']))
print(json.dumps(bad))
"""


@pytest.mark.parametrize("k_init", [0, 1, 2, 3, 4])
def test_golden_find_ac(k_init):
    # the table is built once per process (like the reference), so each k_init gets its own
    code = _FIND_SNIPPET.format(tests=os.path.join(util.ROOT, "tests"), golden=util.GOLDEN)
    r = subprocess.run([sys.executable, "-c", code, str(k_init)], capture_output=True, text=True, check=True)
    assert json.loads(r.stdout.strip().splitlines()[-1]) == []


def test_golden_decode(orc):
    g = load("decode.json")
    for gs in g["streams"]:
        cfg, s = make_golden.synth_stream(gs["ber"])
        pl = util.planted_list(cfg)
        assert len(pl) == gs["n_packets"]
        recs = np.array([util.decode_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), p.clk6, p.uap)
                         for p in pl])
        assert recs[:4].tobytes().hex() == gs["decode_head"]
        assert util.digest(recs) == gs["decode_sha256"]
        hp = "".join(str(orc.orc_header_present(s[p.offset:].ctypes.data, min(3125, len(s) - p.offset))) for p in pl)
        assert hp == gs["header_present"]
        tc = np.array([util.try_clock_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), c)
                       for p in pl[:40] for c in range(64)])
        assert util.digest(tc) == gs["try_clock_sha256"]
        odd = []
        for p in pl[:60]:
            for L in (100, 121, 122, 130, 137, 200, 361, 362, 500):
                odd.append(util.decode_one(orc, "orc", s, p.offset, min(L, len(s) - p.offset), p.clk6, p.uap))
            odd.append(util.decode_one(orc, "orc", s, p.offset, 3125, (p.clk6 + 1) & 63, p.uap))
            odd.append(util.decode_one(orc, "orc", s, p.offset, 3125, p.clk6, (p.uap + 1) & 255))
            odd.append(util.decode_one(orc, "orc", s, p.offset, 3125, p.clk6, p.uap, 0))
        assert util.digest(np.array(odd)) == gs["odd_sha256"]


def noise_type_records(L, prefix):
    g = load("noise_types.json")
    rng = np.random.default_rng(g["seed"])
    recs = []
    for i in range(300):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        n = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        for c in range(0, 64, 7):
            recs.append(util.try_clock_one(L, prefix, sym, 0, n, c))
        recs.append(util.decode_one(L, prefix, sym, 0, n, int(rng.integers(0, 64)), 0))
    return g, np.array(recs)


def test_golden_all_packet_types(orc):
    g, recs = noise_type_records(orc, "orc")
    assert len(recs) == g["count"] and util.digest(recs) == g["sha256"]
    assert set(np.unique(recs["type"]).tolist()) == set(range(16))


# ---- (c) live differential against the compiled reference, where it exists ----
@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/libbtbb_ref.so not built here")
def test_live_reference_decode_differential(orc):
    R = util.ref()
    rng = np.random.default_rng(7)
    cfg = B.synth_cfg(400_000, stride=3500, ber=0.008, seed=77, mix=("DM1", "DH1", "DM3", "FHS", "HV1", "DM5", "DH3"))
    s = B.synth_host(cfg)
    for p in util.planted_list(cfg):
        L = int(rng.choice([3125, p.n_symbols, p.n_symbols - 1, 300]))
        L = min(L, len(s) - p.offset)
        a = util.decode_one(R, "ref", s, p.offset, L, p.clk6, p.uap)
        b = util.decode_one(orc, "orc", s, p.offset, L, p.clk6, p.uap)
        assert a.tobytes() == b.tobytes()
        c = int(rng.integers(0, 64))
        assert util.try_clock_one(R, "ref", s, p.offset, L, c).tobytes() == \
            util.try_clock_one(orc, "orc", s, p.offset, L, c).tobytes()


def test_chain79_fixture(orc):
    """BASELINE configs[2] (79-channel capture): the oracle reproduces the reference's digests for the
    whole chain -- hits, decode with the true clock, raw payload bytes, 64-clock sweep, UAP sieve."""
    g = json.load(open(os.path.join(util.GOLDEN, "chain79.json")))
    assert orc.orc_init(2) == 0
    cfg, s, n = util.chain79_case(g["blocks"])
    hits = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    assert len(hits) == g["hits"] and util.digest(hits) == g["hits_sha256"]
    dec, sv, gs, laps = util.chain79_packets(cfg, hits)
    recs = np.array([util.decode_one(orc, "orc", s, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"])) for p in dec])
    raw = np.array([util.decode_one_raw(orc, "orc", s, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"])) for p in dec])
    tc = np.array([util.try_clock_one(orc, "orc", s, int(p["offset"]), int(p["length"]), c) for p in dec for c in range(64)])
    assert util.digest(recs) == g["decode_sha256"] and util.digest(raw) == g["decode_raw_sha256"]
    assert util.digest(tc) == g["try_clocks_sha256"]
    st, rv = util.sieve_run(orc, "orc", s, sv, gs)
    assert [util.digest(st), util.digest(rv)] == g["sieve_sha256"]
    # the host small-call path (decode_core.h on the CPU) gives the same records
    for i in range(0, len(dec), 5):
        p = dec[i]
        sym = np.ascontiguousarray(s[int(p["offset"]):int(p["offset"]) + int(p["length"])])
        assert B.decode_smallcall(sym, len(sym), int(p["clkn"]), int(p["uap"]))[0].tobytes() == recs[i].tobytes()
        assert B.decode_smallcall(sym, len(sym), mode=B.MODE_TRY_CLOCKS).tobytes() == tc[64 * i:64 * i + 64].tobytes()
    # capture files from the product's host path: records with the raw-payload flag -> the reference's pcap file
    # and pcapng packet blocks, every packet included
    import hashlib
    small = np.array([B.decode_smallcall(np.ascontiguousarray(s[int(p["offset"]):int(p["offset"]) + int(p["length"])]), int(p["length"]),
                                         int(p["clkn"]), int(p["uap"]), mode=B.MODE_FLAG_RAW_PAYLOAD)[0] for p in dec])
    assert small.tobytes() == raw.tobytes()
    meta = util.chain79_meta(dec)
    pcap, png = B.pcap_bredr(hits, small, meta), B.pcapng_bredr_blocks(hits, small, meta)
    assert [len(pcap), hashlib.sha256(pcap).hexdigest()] == g["pcap"]
    assert [len(png), hashlib.sha256(png).hexdigest()] == g["pcapng_blocks"]
