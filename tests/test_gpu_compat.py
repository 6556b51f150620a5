"""The classic btbb_* surface (include/btbb.h) exported by the same shared object behaves
like the reference on the same inputs (first hit, packet fields, decode results)."""
import ctypes as C

import numpy as np
import pytest

import util
from util import B

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("route", ["gpu", "host", "default"])
def test_classic_find_ac_and_decode(product_lib, orc, route):
    """route gpu: every classic call launches kernels; host: every call takes the host small-call path
    (find_ac_host.cpp / decode_host.cpp); default: searches above 8192 positions on the GPU, the rest on the host."""
    L = product_lib
    L.btbb_b200_classic_config(*{"gpu": (-1, 1), "host": (1 << 30, 0), "default": (8192, 0)}[route])
    L.btbb_packet_new.restype = C.c_void_p
    L.btbb_find_ac.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    L.btbb_packet_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint8, C.c_uint32]
    for f in ("btbb_packet_get_lap", "btbb_packet_get_clkn", "btbb_packet_get_header_packed"):
        getattr(L, f).restype = C.c_uint32
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("btbb_packet_get_ac_errors", "btbb_packet_get_type", "btbb_packet_get_lt_addr", "btbb_packet_get_uap",
              "btbb_packet_get_hec", "btbb_packet_get_header_flags"):
        getattr(L, f).restype = C.c_uint8
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("btbb_decode_header", "btbb_decode_payload", "btbb_packet_get_payload_length", "btbb_header_present"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.btbb_packet_set_flag.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.btbb_packet_get_flag.argtypes = [C.c_void_p, C.c_int]
    L.btbb_packet_set_uap.argtypes = [C.c_void_p, C.c_uint8]
    L.btbb_get_payload_packed.argtypes = [C.c_void_p, C.c_void_p]
    L.btbb_packet_unref.argtypes = [C.c_void_p]
    L.try_clock.argtypes = [C.c_int, C.c_void_p]
    L.try_clock.restype = C.c_uint8
    L.crc_check.argtypes = [C.c_int, C.c_void_p]

    assert L.btbb_init(6) == -1 and L.btbb_init(-1) == -1       # bluetooth_packet.c:282-286
    assert L.btbb_init(2) == 0
    assert orc.orc_init(2) == 0
    cfg = B.synth_cfg(300_000, stride=6000, ber=0.001, seed=99, mix=("DM1", "DH1", "DM3", "FHS"))
    s = B.synth_host(cfg)
    n = len(s) - 72
    want = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    pos, seen = 0, 0
    pkt = C.c_void_p(None)
    while seen < 12:
        off = L.btbb_find_ac(s.ctypes.data + pos, n - pos, 0xFFFFFFFF, 2, C.byref(pkt))
        assert off >= 0 and pos + off == want[seen]["offset"]
        assert L.btbb_packet_get_lap(pkt) == want[seen]["lap"]
        assert L.btbb_packet_get_ac_errors(pkt) == want[seen]["ac_errors"]
        assert L.btbb_packet_get_flag(pkt, 0) == 1               # BTBB_WHITENED set by init_packet
        a = pos + off
        p = next((q for q in util.planted_list(cfg) if q.offset == a), None)
        if p is not None:
            avail = min(3125, len(s) - a)
            L.btbb_packet_set_data(pkt, s.ctypes.data + a, avail, 7, p.clk6 << 1)
            assert L.btbb_packet_get_clkn(pkt) == p.clk6
            ref = util.decode_one(orc, "orc", s, a, avail, p.clk6, p.uap)
            L.btbb_packet_set_uap(pkt, p.uap)
            L.btbb_packet_set_flag(pkt, 4, 1)                    # BTBB_CLK6_VALID
            assert L.btbb_decode_header(pkt) == ref["header_ok"]
            assert L.btbb_header_present(pkt) == orc.orc_header_present(s[a:].ctypes.data, avail)
            if ref["header_ok"]:
                assert L.btbb_packet_get_type(pkt) == ref["type"] and L.btbb_packet_get_lt_addr(pkt) == ref["lt_addr"]
                assert L.btbb_packet_get_hec(pkt) == ref["hec"] and L.btbb_packet_get_header_packed(pkt) == ref["header_packed"]
                assert L.btbb_decode_payload(pkt) == ref["rv"]
                assert L.btbb_packet_get_payload_length(pkt) == ref["payload_length"]
                if ref["rv"] >= 2:
                    buf = (C.c_char * 400)()
                    m = L.btbb_get_payload_packed(pkt, buf)
                    assert bytes(buf[:m]) == ref["payload"][:m].tobytes()
                t = util.try_clock_one(orc, "orc", s, a, avail, p.clk6)
                assert L.try_clock(p.clk6, pkt) == t["uap"]
                assert L.crc_check(p.clk6, pkt) == t["rv"]
        pos = a + 1
        seen += 1
    off = L.btbb_find_ac(s.ctypes.data, 3000, 0x9E8B33, 0, C.byref(pkt))
    w = util.find_all(orc, "orc", s, 3000, 0x9E8B33, 0)
    assert off == (int(w[0]["offset"]) if len(w) else -1)
    L.btbb_packet_unref(pkt)
    L.btbb_b200_classic_config(8192, 0)
