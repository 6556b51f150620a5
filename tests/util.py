"""Shared helpers for the test-suite: ctypes views of the oracle (CPU restatement), the
compiled reference (oracle/_ref, optional) and the product library."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from libbtbb_b200 import binding as B  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libbtbb_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "oracle.h", "synth_gen.c")]
        if not os.path.exists(ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
        L = C.CDLL(ORACLE_SO)
        L.orc_syndrome.restype = C.c_uint64
        L.orc_syndrome.argtypes = [C.c_uint64]
        L.orc_gen_syncword.restype = C.c_uint64
        L.orc_gen_syncword.argtypes = [C.c_uint32]
        L.orc_barker_correct.restype = C.c_uint64
        L.orc_find_all.restype = C.c_int64
        L.orc_find_all.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.orc_find_all_mt.restype = C.c_double
        L.orc_find_all_mt.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_fec23.restype = C.c_uint16
        L.orc_crc16.restype = C.c_uint16
        L.orc_hec.restype = C.c_uint8
        L.orc_uap_from_hec.restype = C.c_uint8
        L.orc_table_entries.restype = C.c_long
        L.orc_decode_one.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint8, C.c_int, C.c_void_p]
        L.orc_try_clock_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_decode_one_raw.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint8, C.c_int, C.c_void_p]
        L.orc_header_present.argtypes = [C.c_void_p, C.c_int]
        L.orc_typed_one.restype = None
        L.orc_typed_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_hop_sequence.restype = None
        L.orc_hop_sequence.argtypes = [C.c_uint32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_uap_sieve.restype = None
        L.orc_uap_sieve.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_synth.restype = C.c_int
        L.orc_synth.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The unmodified reference compiled by oracle/Makefile (only where it was built)."""
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_gen_syndrome.restype = C.c_uint64
        L.ref_gen_syndrome.argtypes = [C.c_uint64]
        L.btbb_gen_syncword.restype = C.c_uint64
        L.ref_barker_correct.restype = C.c_uint64
        L.ref_barker_distance.restype = C.c_uint8
        L.ref_find_all.restype = C.c_int64
        L.ref_find_all.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.ref_find_all_mt.restype = C.c_double
        L.ref_find_all_mt.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.ref_fec23.restype = C.c_uint16
        L.ref_crcgen.restype = C.c_uint16
        L.ref_uap_from_hec.restype = C.c_uint8
        L.ref_whitening_bit.restype = C.c_uint8
        L.ref_whitening_index.restype = C.c_uint8
        L.ref_decode_one.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint8, C.c_int, C.c_void_p]
        L.ref_try_clock_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_decode_one_raw.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint8, C.c_int, C.c_void_p]
        L.ref_header_present.argtypes = [C.c_void_p, C.c_int]
        L.ref_typed_one.restype = None
        L.ref_typed_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_hop_sequence.restype = None
        L.ref_hop_sequence.argtypes = [C.c_uint32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.ref_uap_sieve.restype = None
        L.ref_uap_sieve.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _ref = L
    return _ref


def find_all(L, prefix, stream, n, lap, k, cap=1 << 20):
    """All hits of an oracle-like library (prefix 'orc' or 'ref') as a HIT_DTYPE array."""
    hits = np.zeros(cap, dtype=B.HIT_DTYPE)
    cnt = getattr(L, prefix + "_find_all")(stream.ctypes.data, n, lap, k, hits.ctypes.data, cap)
    assert cnt <= cap
    return hits[:cnt].copy()


def decode_one(L, prefix, stream, off, length, clk, uap, whitened=1):
    d = np.zeros(1, dtype=B.DECODED_DTYPE)
    getattr(L, prefix + "_decode_one")(stream[off:].ctypes.data, length, clk, uap, whitened, d.ctypes.data)
    return d[0]


def decode_one_raw(L, prefix, stream, off, length, clk, uap, whitened=1):
    """decode_one with the payload bytes as the decoder left them, whatever rv says."""
    d = np.zeros(1, dtype=B.DECODED_DTYPE)
    getattr(L, prefix + "_decode_one_raw")(stream[off:].ctypes.data, length, clk, uap, whitened, d.ctypes.data)
    return d[0]


def typed_one(L, prefix, sym, length, clock, uap, ptype, fn, whitened=1, raw=0):
    """One forced-type decoder call: fn 0 fhs, 1 DM, 2 DH, 3 EV3, 4 EV4, 5 EV5, 6 HV, -1 btbb_decode_payload, -2 crc_check."""
    d = np.zeros(1, dtype=B.DECODED_DTYPE)
    getattr(L, prefix + "_typed_one")(sym.ctypes.data, length, clock, uap, ptype, whitened, fn, raw, d.ctypes.data)
    return d[0]


def mode_of_fn(fn):
    return B.MODE_CRC_CHECK if fn == -2 else B.MODE_PAYLOAD if fn == -1 else B.MODE_RAW + fn


def forced_type_cases(orc, rng, n):
    """(sym, length, clock, uap, type, fn, whitened): noise and FEC-clean packets for the single decoders,
    the type field forced -- matching the decoder or not."""
    own = {0: [2], 1: [3, 8, 10, 14], 2: [4, 9, 11, 15], 3: [7], 4: [12], 5: [13], 6: [5, 6, 7]}
    for i in range(n):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        if i % 5 == 4:
            sym[68 + 3 * int(rng.integers(0, 18))] ^= 1
        if i % 2 == 0:
            for b, d in enumerate(rng.integers(0, 1024, 183)):
                cw = orc.orc_fec23(int(d))
                sym[122 + 15 * b:122 + 15 * b + 15] = [(cw >> t) & 1 for t in range(15)]
            if i % 6 == 0:
                b = int(rng.integers(0, 40))
                sym[122 + 15 * b:122 + 15 * b + 15] ^= 1
                sym[122 + 15 * b + 3] ^= 1
        length = int(rng.choice([3125, 3125, 1500, 700, 362, 361, 250, 140, 122, 100]))
        fn = int(rng.integers(-2, 7))
        ptype = int(rng.choice(own[fn])) if fn >= 0 and rng.integers(0, 3) else int(rng.integers(0, 16))
        yield sym, length, int(rng.integers(0, 64)), int(rng.integers(0, 256)), ptype, fn, int(rng.integers(0, 8) != 0)


def try_clock_one(L, prefix, stream, off, length, clock, whitened=1):
    d = np.zeros(1, dtype=B.DECODED_DTYPE)
    getattr(L, prefix + "_try_clock_one")(stream[off:].ctypes.data, length, clock, whitened, d.ctypes.data)
    return d[0]


def planted_list(cfg):
    n_slots = (cfg.first_symbol + cfg.n_symbols + cfg.stride - 1) // cfg.stride
    out = []
    for s in range(cfg.first_symbol // cfg.stride, n_slots):
        p = B.planted(cfg, s)
        if p.offset >= cfg.first_symbol and p.offset + p.n_symbols <= cfg.first_symbol + cfg.n_symbols:
            out.append(p)
    return out


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def plant_syncwords(stream, rng, count, max_errors, laps=None):
    """Overwrite `count` random places with sync words carrying 0..max_errors bit flips."""
    O = oracle()
    placed = []
    n = len(stream) - 64
    for _ in range(count):
        lap = int(rng.choice(laps)) if laps is not None else int(rng.integers(0, 1 << 24))
        sw = O.orc_gen_syncword(lap)
        ne = int(rng.integers(0, max_errors + 1))
        for e in rng.choice(64, ne, replace=False):
            sw ^= 1 << int(e)
        p = int(rng.integers(0, n))
        stream[p:p + 64] = [(sw >> i) & 1 for i in range(64)]
        placed.append((p, lap, ne))
    return placed


def sieve_case(n_slots=600, stride=4000, n_laps=12, ber=0.0, seed=B.DEFAULT_SEED, mix=("ID", "DM1", "DH1", "DM3", "FHS", "HV1"),
               clk_step=1, coherent=True, repeat_first=0):
    """A piconet-coherent synthetic capture plus the sieve's inputs: the stream, the packets
    grouped by LAP in arrival order (CLKN = slot * clk_step, channel = slot % 79), the group
    boundaries and the ground truth {lap: (uap, clk6 of slot 0)}."""
    n = n_slots * stride
    cfg = B.synth_cfg(n, stride=stride, n_laps=n_laps, ber=ber, mix=mix, seed=seed, piconets=coherent)
    stream = B.synth_host(cfg)
    by_lap, truth = {}, {}
    for slot in range(n_slots):
        p = B.planted(cfg, slot)
        if p.offset + p.n_symbols > n:
            continue
        by_lap.setdefault(p.lap, []).append((slot, p))
        truth[p.lap] = (p.uap, (p.clk6 - slot) & 63)
    laps = sorted(by_lap)
    pkts = np.zeros(sum(len(v) for v in by_lap.values()), dtype=B.PKTIN_DTYPE)
    gs, i = [0], 0
    for lap in laps:
        for slot, p in by_lap[lap]:
            pkts[i]["offset"], pkts[i]["length"] = p.offset, min(3125, n - p.offset)
            pkts[i]["clkn"], pkts[i]["whitened"], pkts[i]["reserved"] = (slot * clk_step) & 0xFFFFFFFF, 1, slot % 79
            i += 1
        gs.append(i)
    if repeat_first:
        # every piconet sees its first header-bearing packet again and again at the same CLKN: no
        # candidate is ever eliminated, which runs into the reference's 1000-packet limit
        rep, gs2 = [], [0]
        for g in range(len(laps)):
            grp = pkts[gs[g]:gs[g + 1]]
            pick = next((q for q in grp if q["length"] >= 400), grp[0])
            rep.append(np.repeat(pick[None], repeat_first))
            gs2.append(gs2[-1] + repeat_first)
        pkts, gs = np.concatenate(rep), gs2
    return stream, pkts, np.array(gs, dtype=np.int64), laps, truth


def sieve_run(L, prefix, stream, pkts, gs, states=None):
    """Run an oracle-like library's sieve; returns (states, rv)."""
    st = np.zeros(len(gs) - 1, dtype=B.SIEVE_DTYPE) if states is None else states.copy()
    rv = np.zeros(len(pkts), dtype=np.int8)
    getattr(L, prefix + "_uap_sieve")(stream.ctypes.data, len(stream), pkts.ctypes.data, len(pkts),
                                      gs.ctypes.data, len(st), st.ctypes.data, rv.ctypes.data)
    return st, rv



# ---- crafted packets: the CRC-success paths of every search (EV3 / EV4 / EV5 lengths, fhs clocks, DV) ----
def _bits(v, n):
    return [(v >> i) & 1 for i in range(n)]


def _whiten(orc, bits, clk, skip):
    return [int(b) ^ orc.orc_whiten_bit(clk, skip + i) for i, b in enumerate(bits)]


def _fec23(orc, bits):
    out = []
    bits = list(bits) + [0] * (-len(bits) % 10)
    for b in range(0, len(bits), 10):
        d = int(sum(int(bits[b + i]) << i for i in range(10)))
        out += _bits(orc.orc_fec23(d), 15)
    return out


def _with_crc(orc, body_bytes, uap):
    bits = [b for byte in body_bytes for b in _bits(int(byte), 8)]
    arr = np.array(bits, dtype=np.uint8)
    crc = orc.orc_crc16(arr.ctypes.data, len(bits), uap)
    return bits + _bits(crc, 16)


def _craft(orc, rng, ptype, clk, uap, payload_symbols):
    sym = rng.integers(0, 2, 3125, dtype=np.uint8)
    d10 = int(rng.integers(0, 8)) | (ptype << 3) | (int(rng.integers(0, 8)) << 7)
    hdr = _bits(d10 | (orc.orc_hec(d10, uap) << 10), 18)
    sym[68:122] = np.repeat(np.array(_whiten(orc, hdr, clk, 0), dtype=np.uint8), 3)
    n = min(len(payload_symbols), 3125 - 122)
    sym[122:122 + n] = payload_symbols[:n]
    return sym


def crafted_packets(orc, rng, count):
    """Packets whose payload CRC closes on the paths a random capture hardly ever reaches: EV4 at a
    chosen length (optionally in front of an uncorrectable block / a short capture), FHS whitened
    with its own clock or with one of 32..63, DV, EV3 / EV5.  Yields (kind, symbols, length, clk, uap)."""
    for i in range(count):
        clk, uap = int(rng.integers(0, 64)), int(rng.integers(0, 256))
        kind = i % 4
        if kind == 0:
            L = int(rng.integers(3, 120))
            pay = _with_crc(orc, rng.integers(0, 256, L - 2), uap)
            pay += list(rng.integers(0, 2, 10 * 98 - len(pay) if 10 * 98 > len(pay) else 0))
            sym = _craft(orc, rng, 12, clk, uap, _fec23(orc, _whiten(orc, pay, clk, 18)))
            if i % 8 == 4:
                b = int(rng.integers(0, 90))
                sym[122 + 15 * b:122 + 15 * b + 3] ^= 1
            n = int(rng.choice([3125, 122 + 15 * int(rng.integers(1, 98)) + int(rng.integers(0, 15))]))
        elif kind == 1:
            pay = _with_crc(orc, rng.integers(0, 256, 18), uap)
            clk2 = clk if i % 8 == 1 else int(rng.integers(32, 64))
            sym = _craft(orc, rng, 2, clk, uap, _fec23(orc, _whiten(orc, pay, clk2, 18)))
            if i % 16 == 5:
                sym[122 + int(rng.integers(0, 240))] ^= 1
            n = 3125 if i % 5 else 362
        elif kind == 2:
            nb = int(rng.integers(0, 10))
            hdr = int(rng.integers(0, 8)) | (nb << 3)
            pay = _with_crc(orc, [hdr] + list(rng.integers(0, 256, nb)), uap)
            s = list(rng.integers(0, 2, 80)) + _fec23(orc, _whiten(orc, pay, clk, 18))
            sym = _craft(orc, rng, 8, clk, uap, s)
            n = 3125 if i % 3 else 202 + 15 * ((8 * (nb + 3) + 9) // 10) + int(rng.integers(0, 40))
        else:
            sym = _craft(orc, rng, 7 if i % 8 == 3 else 13, clk, uap, rng.integers(0, 2, 3000, dtype=np.uint8))
            n = int(rng.choice([3125, 1000, 400, 200]))
        yield ("ev4", "fhs_own" if i % 8 == 1 else "fhs_other", "dv", "ev35")[kind] if kind != 1 else ("fhs_own" if i % 8 == 1 else "fhs_other"), sym, n, clk, uap


def ev35_hunt(orc, rng, count):
    """EV3 / EV5 packets on random symbols (a CRC closure is luck: about one in 400 for EV5)."""
    for i in range(count):
        clk, uap = int(rng.integers(0, 64)), int(rng.integers(0, 256))
        yield _craft(orc, rng, 13 if i % 4 else 7, clk, uap, rng.integers(0, 2, 3000, dtype=np.uint8)), clk, uap


# ---- BASELINE configs[2]: the full chain on a 79-channel interleaved capture (SURVEY.md 8d cfg 3) ----
CHAIN_BLK, CHAIN_CH = 4096, 79


def chain79_case(blocks=6, ber=0.004, seed=B.DEFAULT_SEED):
    """Capture laid out [block][79 channels][4096 symbols], every channel block carrying one planted
    packet of a piconet-coherent capture (one UAP per LAP, CLK1-6 advancing with the slot)."""
    n = blocks * CHAIN_CH * CHAIN_BLK
    cfg = B.synth_cfg(n + 63, stride=CHAIN_BLK, n_laps=8, ber=ber, seed=seed, mix=("DM1", "DM3", "DH1", "FHS", "HV1"), piconets=True)
    return cfg, B.synth_host(cfg), n


def chain79_packets(cfg, hits):
    """Per hit what the caller of the chain knows: the symbols left in the hit's channel block, the
    slot as CLKN, slot % 79 as channel; for hits that are planted packets the true CLK1-6 / UAP
    (else 0 / 0).  Returns (pkt_in for decode with the true clock, pkt_in for the sieve)."""
    dec = np.zeros(len(hits), dtype=B.PKTIN_DTYPE)
    for i, h in enumerate(hits):
        off = int(h["offset"])
        slot = off // CHAIN_BLK
        p = B.planted(cfg, slot)
        dec[i]["offset"], dec[i]["length"], dec[i]["whitened"] = off, min(3125, (slot + 1) * CHAIN_BLK - off), 1
        if p.offset == off:
            dec[i]["clkn"], dec[i]["uap"] = p.clk6, p.uap
        dec[i]["reserved"] = slot % CHAIN_CH
    order, gs, laps = B.group_by_lap(hits)
    sv = dec[order].copy()
    sv["clkn"] = sv["offset"] // CHAIN_BLK
    sv["uap"] = 0
    return dec, sv, gs, laps


def chain79_meta(dec):
    """capture metadata for the pcap writers: 625 us slots, channel from the packet record"""
    meta = np.zeros(len(dec), dtype=B.PCAP_META_DTYPE)
    for i, p in enumerate(dec):
        meta[i]["ns"] = 1_700_000_000_000_000_000 + 625_000 * (int(p["offset"]) // CHAIN_BLK)
        meta[i]["sigdbm"], meta[i]["noisedbm"] = -30 - i % 40, -95 + i % 7
        meta[i]["channel"], meta[i]["transport"], meta[i]["modulation"] = int(p["reserved"]) & 0xff, 1, 0
    return meta
