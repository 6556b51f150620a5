"""BR/EDR pcap emission: btbb_b200_pcap_file_header / _bredr_records against
the reference's own btbb_pcap_create_file / btbb_pcap_append_packet (oracle/_ref) and against a
committed fixture of its output.  The serialiser is host code, so this runs without a GPU: the
decoded records come from the oracle's restatement of btbb_decode (itself pinned to the reference).
Packets whose payload decode failed are left out -- the reference logs the bytes its decoder left
behind for them, which btbb_b200_decoded does not carry (include/btbb_b200.h)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import util
from util import B

FIXTURE = os.path.join(util.GOLDEN, "pcap.json")
CASES = {"clean": dict(ber=0.0, seed=5), "ber_0.3pct": dict(ber=0.003, seed=6)}


def build_case(ber, seed, n_slots=160, stride=5000):
    n = n_slots * stride
    cfg = B.synth_cfg(n, stride=stride, n_laps=8, ber=ber, seed=seed, mix=("ID", "DM1", "DH1", "DM3", "FHS", "HV1", "DH3"))
    stream = B.synth_host(cfg)
    planted = [p for p in util.planted_list(cfg)]
    rng = np.random.default_rng(seed)
    hits = np.zeros(len(planted), dtype=B.HIT_DTYPE)
    pkts = np.zeros(len(planted), dtype=B.PKTIN_DTYPE)
    meta = np.zeros(len(planted), dtype=B.PCAP_META_DTYPE)
    for i, p in enumerate(planted):
        hits[i]["offset"], hits[i]["lap"], hits[i]["ac_errors"] = p.offset, p.lap, rng.integers(0, 3)
        pkts[i]["offset"], pkts[i]["length"] = p.offset, min(3125, n - p.offset)
        pkts[i]["clkn"], pkts[i]["uap"], pkts[i]["whitened"] = p.clk6 + 64 * int(rng.integers(0, 1000)), p.uap, 1
        meta[i]["ns"] = 1_700_000_000_000_000_000 + 625_000 * i + int(rng.integers(0, 1000))
        meta[i]["sigdbm"], meta[i]["noisedbm"] = -int(rng.integers(20, 90)), -int(rng.integers(20, 100))
        meta[i]["channel"], meta[i]["transport"], meta[i]["modulation"] = rng.integers(0, 79), rng.integers(0, 5), rng.integers(0, 3)
    return stream, hits, pkts, meta


def oracle_records(stream, pkts):
    O = util.oracle()
    return np.array([util.decode_one(O, "orc", stream, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"]))
                     for p in pkts])


def well_defined(dec):
    return (dec["rv"] >= 2) | (dec["payload_length"] == 0)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("ref_id", [(B.LAP_ANY, 0xFF), (0x9E8B33, 0x42)])
def test_pcap_records_match_reference_and_fixture(name, ref_id, tmp_path):
    stream, hits, pkts, meta = build_case(**CASES[name])
    dec = oracle_records(stream, pkts)
    keep = well_defined(dec)
    assert keep.sum() >= 0.8 * len(keep) and (dec["rv"] >= 10).sum() > 20 and (dec["payload_length"] == 0).sum() > 5
    reflap, refuap = ref_id
    got = B.pcap_bredr(hits[keep], dec[keep], meta[keep], reflap, refuap)
    fx = json.load(open(FIXTURE))[f"{name}/{reflap:x}/{refuap:x}"]
    assert [int(keep.sum()), len(got)] == fx["shape"]
    assert hashlib.sha256(got).hexdigest() == fx["sha256"]
    if util.have_ref():
        R = util.ref()
        path = str(tmp_path / "ref.pcap").encode()
        h, p, m = hits[keep].copy(), pkts[keep].copy(), meta[keep].copy()
        rv = np.zeros(len(h), dtype=np.int32)
        R.ref_pcap_bredr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_uint32, C.c_uint8, C.c_void_p]
        assert R.ref_pcap_bredr(path, stream.ctypes.data, len(stream), h.ctypes.data, p.ctypes.data, m.ctypes.data, len(h),
                                reflap, refuap, rv.ctypes.data) == 0
        want = open(path.decode(), "rb").read()
        assert got == want


def test_pcap_size_query_and_small_buffer(product_lib):
    stream, hits, pkts, meta = build_case(**CASES["clean"])
    dec = oracle_records(stream, pkts)
    L = B.lib()
    need = L.btbb_b200_pcap_bredr_records(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), B.LAP_ANY, 0xFF, None, 0)
    assert need == sum(16 + 22 + min(400, int(d["payload_length"])) for d in dec)
    small = np.full(need - 1, 0xAB, dtype=np.uint8)
    assert L.btbb_b200_pcap_bredr_records(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), B.LAP_ANY, 0xFF,
                                          small.ctypes.data, need - 1) == need
    assert (small == 0xAB).all()                       # nothing written when it does not fit
    assert L.btbb_b200_pcap_file_header(small.ctypes.data, 23) == -1


def _ref_pcap(tmp_path, stream, hits, pkts, meta, reflap=B.LAP_ANY, refuap=0xFF):
    R = util.ref()
    path = str(tmp_path / "ref_all.pcap").encode()
    rv = np.zeros(len(hits), dtype=np.int32)
    R.ref_pcap_bredr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_uint32, C.c_uint8, C.c_void_p]
    assert R.ref_pcap_bredr(path, stream.ctypes.data, len(stream), hits.ctypes.data, pkts.ctypes.data, meta.ctypes.data,
                            len(hits), reflap, refuap, rv.ctypes.data) == 0
    return open(path.decode(), "rb").read()


@pytest.mark.parametrize("ber,seed", [(0.003, 6), (0.01, 7), (0.03, 8)])
def test_raw_payload_semantics_and_full_file_parity(ber, seed, tmp_path):
    """What separates the serialiser from byte-exact parity on EVERY packet is only the payload of
    failed decodes: with records that carry pkt->payload as the reference's decoders leave it
    (orc_decode_one_raw, checked here against the reference itself) the whole file matches."""
    if not util.have_ref():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    O, R = util.oracle(), util.ref()
    stream, hits, pkts, meta = build_case(ber=ber, seed=seed)
    raw = np.array([util.decode_one_raw(O, "orc", stream, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"]))
                    for p in pkts])
    ref_raw = np.array([util.decode_one_raw(R, "ref", stream, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"]))
                        for p in pkts])
    assert raw.tobytes() == ref_raw.tobytes()
    assert ((raw["rv"] < 2) & (raw["payload_length"] > 0)).sum() > 0        # the case in question occurs
    assert B.pcap_bredr(hits, raw, meta) == _ref_pcap(tmp_path, stream, hits, pkts, meta)


def _epbs(data):
    """the enhanced packet blocks of a pcapng file, pad bytes zeroed (upstream leaves stack garbage there)"""
    out, at = bytearray(), 0
    while at + 12 <= len(data):
        btype = int.from_bytes(data[at:at + 4], "little")
        blen = int.from_bytes(data[at + 4:at + 8], "little")
        if blen < 12 or at + blen > len(data):
            break
        if btype == 6:
            blk = bytearray(data[at:at + blen])
            cap = int.from_bytes(blk[20:24], "little")
            for i in range(28 + cap, blen - 8):
                blk[i] = 0
            out += blk
        at += blen
    return bytes(out)


@pytest.mark.parametrize("ber,seed", [(0.0, 5), (0.01, 7)])
def test_pcapng_blocks_match_reference(ber, seed, tmp_path):
    """btbb_b200_pcapng_bredr_blocks == the enhanced packet blocks btbb_pcapng_append_packet writes
    (pcapng-bt.c:176-264), for every packet when the records carry the raw payload bytes."""
    if not util.have_ref():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    O, R = util.oracle(), util.ref()
    stream, hits, pkts, meta = build_case(ber=ber, seed=seed)
    raw = np.array([util.decode_one_raw(O, "orc", stream, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"]))
                    for p in pkts])
    R.ref_pcapng_bredr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_uint32, C.c_uint8, C.c_void_p]
    for reflap, refuap in ((B.LAP_ANY, 0xFF), (0x9E8B33, 0x42)):
        path = str(tmp_path / f"ref_{reflap:x}.pcapng").encode()
        rv = np.zeros(len(hits), dtype=np.int32)
        assert R.ref_pcapng_bredr(path, stream.ctypes.data, len(stream), hits.ctypes.data, pkts.ctypes.data, meta.ctypes.data,
                                  len(hits), reflap, refuap, rv.ctypes.data) == 0
        want = _epbs(open(path.decode(), "rb").read())
        got = B.pcapng_bredr_blocks(hits, raw, meta, reflap, refuap)
        assert len(want) > 36 * len(hits) and got == want
