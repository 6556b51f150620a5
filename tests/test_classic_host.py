"""The classic single-packet calls (include/btbb.h: btbb_decode_header / btbb_decode_payload /
try_clock / crc_check / fhs / DM / ... / btbb_header_present) on their default route, the host
small-call path (decode_host.cpp = decode_core.h compiled for the CPU), against the oracle -- and,
where oracle/_ref holds it, against the unmodified reference driven through the same calls."""
import ctypes as C

import numpy as np
import pytest

import util
from util import B

BTBB_WHITENED, BTBB_UAP_VALID, BTBB_CLK6_VALID, BTBB_HAS_PAYLOAD = 0, 2, 4, 7


def _proto(L):
    L.btbb_packet_new.restype = C.c_void_p
    L.btbb_packet_unref.argtypes = [C.c_void_p]
    L.btbb_packet_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint8, C.c_uint32]
    L.btbb_packet_set_flag.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.btbb_packet_get_flag.argtypes = [C.c_void_p, C.c_int]
    L.btbb_packet_set_uap.argtypes = [C.c_void_p, C.c_uint8]
    for f in ("btbb_packet_get_header_packed",):
        getattr(L, f).restype = C.c_uint32
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("btbb_packet_get_type", "btbb_packet_get_lt_addr", "btbb_packet_get_uap", "btbb_packet_get_hec",
              "btbb_packet_get_header_flags"):
        getattr(L, f).restype = C.c_uint8
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("btbb_decode_header", "btbb_decode_payload", "btbb_packet_get_payload_length", "btbb_header_present"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.btbb_get_payload_packed.argtypes = [C.c_void_p, C.c_void_p]
    L.try_clock.argtypes = [C.c_int, C.c_void_p]
    L.try_clock.restype = C.c_uint8
    for f in ("crc_check", "fhs", "DM", "DH", "EV3", "EV4", "EV5", "HV"):
        getattr(L, f).argtypes = [C.c_int, C.c_void_p]
    return L


def _drive(L, sym, n, clk, uap):
    """What a caller of the classic API sees for one packet: (header_ok, header_packed, type, lt_addr,
    hec, rv, payload_length, has_payload, payload bytes, [try_clock uap, crc_check rv] for 4 clocks)."""
    pkt = L.btbb_packet_new()
    L.btbb_packet_set_flag(pkt, BTBB_WHITENED, 1)
    L.btbb_packet_set_data(pkt, sym.ctypes.data, n, 3, clk << 1)
    L.btbb_packet_set_uap(pkt, uap)
    L.btbb_packet_set_flag(pkt, BTBB_CLK6_VALID, 1)
    out = [L.btbb_header_present(pkt)]
    ok = L.btbb_decode_header(pkt)
    out += [ok, L.btbb_packet_get_header_packed(pkt)]
    if ok:
        out += [L.btbb_packet_get_type(pkt), L.btbb_packet_get_lt_addr(pkt), L.btbb_packet_get_hec(pkt), L.btbb_packet_get_header_flags(pkt)]
        rv = L.btbb_decode_payload(pkt)
        buf = (C.c_char * 400)()
        m = L.btbb_get_payload_packed(pkt, buf)
        out += [rv, m, L.btbb_packet_get_flag(pkt, BTBB_HAS_PAYLOAD), bytes(buf[:max(m, 0)])]
    for c in (clk, (clk + 1) & 63, (clk + 33) & 63, 5):
        p2 = L.btbb_packet_new()
        L.btbb_packet_set_flag(p2, BTBB_WHITENED, 1)
        L.btbb_packet_set_data(p2, sym.ctypes.data, n, 3, 0)
        u = L.try_clock(c, p2)
        out += [u, L.btbb_packet_get_type(p2), L.crc_check(c, p2), L.btbb_packet_get_payload_length(p2)]
        L.btbb_packet_unref(p2)
    L.btbb_packet_unref(pkt)
    return out


def test_classic_packet_calls_on_the_host_route(product_lib, orc):
    L = _proto(product_lib)
    L.btbb_b200_classic_config(8192, 0)
    rng = np.random.default_rng(11)
    cases = [(sym, n, clk, uap) for _, sym, n, clk, uap in util.crafted_packets(orc, rng, 120)]
    cfg = B.synth_cfg(400_000, stride=3400, ber=0.004, seed=5151, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    for p in util.planted_list(cfg):
        n = min(3125, len(s) - p.offset)
        cases.append((np.ascontiguousarray(s[p.offset:p.offset + n]), n, p.clk6, p.uap))
    R = _proto(C.CDLL(util.REF_SO)) if util.have_ref() else None
    if R is not None:
        R.btbb_init(2)
    for i, (sym, n, clk, uap) in enumerate(cases):
        got = _drive(L, sym, n, clk, uap)
        want = util.decode_one(orc, "orc", sym, 0, n, clk, uap)
        assert got[0] == orc.orc_header_present(sym.ctypes.data, n)
        assert got[1] == want["header_ok"]
        if want["header_ok"]:
            assert got[2] == want["header_packed"] and got[3:7] == [want["type"], want["lt_addr"], want["hec"], want["flags"]]
            assert got[7] == want["rv"] and got[8] == want["payload_length"] and got[9] == 1
            if want["rv"] >= 2:
                assert got[10] == want["payload"][:got[8]].tobytes()
        if R is not None:      # the reference itself, same calls (it prints nothing on these paths)
            assert got == _drive(R, sym, n, clk, uap), i
