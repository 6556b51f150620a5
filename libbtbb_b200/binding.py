"""ctypes view of the C ABI in include/btbb_b200.h (tests and bench.py use this; the
product itself is the shared object).  Importing never falls back to anything: if
libbtbb.so.1 is missing the import raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbtbb.so.1")
LAP_ANY = 0xFFFFFFFF


class Hit(C.Structure):
    _fields_ = [("offset", C.c_int64), ("lap", C.c_uint32), ("ac_errors", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Decoded(C.Structure):
    _fields_ = [("header_ok", C.c_int32), ("rv", C.c_int32), ("uap", C.c_uint8), ("type", C.c_uint8),
                ("lt_addr", C.c_uint8), ("flags", C.c_uint8), ("hec", C.c_uint8), ("llid", C.c_uint8),
                ("flow", C.c_uint8), ("has_payload", C.c_uint8), ("payload_header_length", C.c_int32),
                ("payload_length", C.c_int32), ("header_packed", C.c_uint32), ("payload", C.c_uint8 * 344)]


class PktIn(C.Structure):
    _fields_ = [("offset", C.c_int64), ("length", C.c_int32), ("clkn", C.c_uint32), ("uap", C.c_uint8),
                ("whitened", C.c_uint8), ("type", C.c_uint8), ("pad", C.c_uint8), ("reserved", C.c_uint32)]


class HopCfg(C.Structure):
    _fields_ = [("address", C.c_uint32), ("afh", C.c_uint8), ("aliased", C.c_uint8), ("pad", C.c_uint8 * 2),
                ("afh_map", C.c_uint8 * 10), ("pad2", C.c_uint8 * 2)]


class SynthCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_symbols", C.c_int64), ("first_symbol", C.c_int64),
                ("stride", C.c_int32), ("n_laps", C.c_int32), ("ber_q32", C.c_uint32),
                ("packet_mix", C.c_uint32), ("fixed_lap", C.c_uint32), ("reserved", C.c_uint32)]


class Planted(C.Structure):
    _fields_ = [("offset", C.c_int64), ("lap", C.c_uint32), ("uap", C.c_uint8), ("kind", C.c_uint8),
                ("clk6", C.c_uint8), ("lt_addr", C.c_uint8), ("n_symbols", C.c_int32), ("body_bytes", C.c_int32)]


HIT_DTYPE = np.dtype([("offset", "<i8"), ("lap", "<u4"), ("ac_errors", "u1"), ("pad", "u1", (3,))])
DECODED_DTYPE = np.dtype([("header_ok", "<i4"), ("rv", "<i4"), ("uap", "u1"), ("type", "u1"), ("lt_addr", "u1"),
                          ("flags", "u1"), ("hec", "u1"), ("llid", "u1"), ("flow", "u1"), ("has_payload", "u1"),
                          ("payload_header_length", "<i4"), ("payload_length", "<i4"), ("header_packed", "<u4"),
                          ("payload", "u1", (344,))])
PKTIN_DTYPE = np.dtype([("offset", "<i8"), ("length", "<i4"), ("clkn", "<u4"), ("uap", "u1"),
                        ("whitened", "u1"), ("type", "u1"), ("pad", "u1"), ("reserved", "<u4")])
SIEVE_DTYPE = np.dtype([("flags", "<u4"), ("first_pkt_time", "<u4"), ("clk_offset", "<i4"), ("packets_observed", "<i4"),
                        ("total_packets_observed", "<i4"), ("uap", "u1"), ("used_channels", "u1"), ("afh_map", "u1", (10,)),
                        ("clock6_candidates", "<i2", (64,))])
PCAP_META_DTYPE = np.dtype([("ns", "<u8"), ("sigdbm", "i1"), ("noisedbm", "i1"), ("channel", "u1"), ("transport", "u1"),
                            ("modulation", "u1"), ("pad", "u1", (3,))])
assert PCAP_META_DTYPE.itemsize == 16
SIEVE_NOT_CALLED = -2
F_UAP_VALID, F_CLK6_VALID, F_GOT_FIRST_PACKET = 1 << 2, 1 << 4, 1 << 10
assert SIEVE_DTYPE.itemsize == 160
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 16
assert DECODED_DTYPE.itemsize == C.sizeof(Decoded) == 372
assert PKTIN_DTYPE.itemsize == C.sizeof(PktIn) == 24

KIND = {"ID": 0, "DM1": 1, "DH1": 2, "DM3": 3, "FHS": 4, "HV1": 5, "DM5": 6, "DH3": 7}
DEFAULT_SEED = 0xB200B7BB

_vp, _i64, _u32, _int = C.c_void_p, C.c_int64, C.c_uint32, C.c_int
_PROTOS = {
    "btbb_b200_create": (_int, [_int, _int, C.POINTER(_vp)]),
    "btbb_b200_destroy": (None, [_vp]),
    "btbb_b200_device": (_int, [_vp]),
    "btbb_b200_table_errors": (_int, [_vp]),
    "btbb_b200_last_error": (C.c_char_p, []),
    "btbb_b200_find_ac_dev": (_int, [_vp, _vp, _i64, _u32, _int, _vp, _i64, C.POINTER(_i64), _vp]),
    "btbb_b200_find_ac_dev_begin": (_int, [_vp, _vp, _i64, _u32, _int, _vp, _i64, _vp]),
    "btbb_b200_find_ac_dev_end": (_int, [_vp, C.POINTER(_i64)]),
    "btbb_b200_set_offset_bias": (_int, [_vp, _i64]),
    "btbb_b200_set_option": (_int, [_vp, _int, _i64]),
    "btbb_b200_set_profiling": (_int, [_vp, _int]),
    "btbb_b200_last_scan_kernel_ms": (_int, [_vp, C.POINTER(C.c_float)]),
    "btbb_b200_find_ac_packed_dev": (_int, [_vp, _vp, _i64, _u32, _int, _vp, _i64, C.POINTER(_i64), _vp]),
    "btbb_b200_find_ac_enqueue": (_int, [_vp, _vp, _i64, _u32, _int, _vp, _i64, _vp, _vp]),
    "btbb_b200_find_ac_host": (_int, [_vp, _vp, _i64, _u32, _int, _vp, _i64, C.POINTER(_i64)]),
    "btbb_b200_decode_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "btbb_b200_decode_host": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp]),
    "btbb_b200_try_clocks_compact_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "btbb_b200_decode_smallcall": (_int, [_vp, _int, _u32, C.c_uint8, _int, C.c_uint8, _int, _vp]),
    "btbb_b200_shard_unique_id": (_int, [_vp]),
    "btbb_b200_shard_init": (_int, [_vp, _vp, _int, _int, _i64, _int]),
    "btbb_b200_shard_info": (_int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "btbb_b200_shard_destroy": (_int, [_vp]),
    "btbb_b200_find_ac_sharded_begin": (_int, [_vp, _vp, _i64, _i64, _u32, _int, _vp]),
    "btbb_b200_find_ac_sharded_end": (_int, [_vp, C.POINTER(_i64)]),
    "btbb_b200_find_ac_sharded_next": (_int, [_vp, _vp, _i64, _i64, _u32, _int, _vp, C.POINTER(_i64)]),
    "btbb_b200_find_ac_sharded_gather": (_int, [_vp, C.POINTER(_vp), C.POINTER(_i64), _vp, C.POINTER(_i64)]),
    "btbb_b200_find_ac_sharded_dev": (_int, [_vp, _vp, _i64, _i64, _u32, _int, _vp, _i64, _vp, C.POINTER(_i64), _vp]),
    "btbb_b200_header_present_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "btbb_b200_header_present_host": (_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "btbb_b200_classic_config": (None, [_int, _int]),
    "btbb_b200_uap_sieve_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "btbb_b200_uap_sieve_host": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "btbb_b200_group_by_lap": (_i64, [_vp, _i64, _vp, _vp, _vp]),
    "btbb_b200_hop_sequence_dev": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "btbb_b200_hop_winnow": (_int, [_vp, _vp, _u32, _int, _vp, _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "btbb_b200_pcap_file_header": (_i64, [_vp, _i64]),
    "btbb_b200_pcap_bredr_records": (_i64, [_vp, _vp, _vp, _i64, _u32, C.c_uint8, _vp, _i64]),
    "btbb_b200_pcapng_bredr_blocks": (_i64, [_vp, _vp, _vp, _i64, _u32, C.c_uint8, _vp, _i64]),
    "btbb_b200_find_first_smallcall": (_int, [_vp, _int, _u32, _int, _int, _vp, C.POINTER(_int)]),
    "btbb_b200_capture_records_dev": (_int, [_vp, _int, _vp, _vp, _vp, _i64, _u32, C.c_uint8, _vp, _i64, C.POINTER(_i64), _vp]),
    "btbb_b200_synth_host": (_int, [C.POINTER(SynthCfg), _vp]),
    "btbb_b200_synth_dev": (_int, [C.POINTER(SynthCfg), _vp, _vp]),
    "btbb_b200_synth_planted": (_int, [C.POINTER(SynthCfg), _i64, C.POINTER(Planted)]),
}
# the classic surface (include/btbb.h) -- checked for presence by tests/test_abi.py
CLASSIC_SYMBOLS = [
    "btbb_init", "btbb_get_release", "btbb_get_version", "btbb_packet_new", "btbb_packet_ref",
    "btbb_packet_unref", "btbb_find_ac", "btbb_packet_set_flag", "btbb_packet_get_flag",
    "btbb_packet_get_lap", "btbb_packet_set_uap", "btbb_packet_get_uap", "btbb_packet_get_nap",
    "btbb_packet_set_modulation", "btbb_packet_set_transport", "btbb_packet_get_modulation",
    "btbb_packet_get_transport", "btbb_packet_get_channel", "btbb_packet_get_ac_errors",
    "btbb_packet_get_clkn", "btbb_packet_get_header_packed", "btbb_packet_set_data", "btbb_get_symbols",
    "btbb_packet_get_payload_length", "btbb_get_payload", "btbb_get_payload_packed", "btbb_packet_get_type",
    "btbb_packet_get_lt_addr", "btbb_packet_get_header_flags", "btbb_packet_get_hec", "btbb_gen_syncword",
    "btbb_decode_header", "btbb_decode_payload", "btbb_decode", "btbb_print_packet", "btbb_header_present",
    # exported helpers the reference's piconet layer links against (bluetooth_packet.h:115-144)
    "try_clock", "crc_check", "fhs", "DM", "DH", "EV3", "EV4", "EV5", "HV", "tun_format",
    "lap_from_fhs", "uap_from_fhs", "nap_from_fhs", "clock_from_fhs", "find_known_lap",
    "promiscuous_packet_search",
]

_lib = None


def lib():
    """Load libbtbb.so.1; raises if it has not been built (python -m libbtbb_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m libbtbb_b200.build` (there is no fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L.btbb_gen_syncword.restype = C.c_uint64
        L.btbb_gen_syncword.argtypes = [_int]
        _lib = L
    return _lib


class BtbbError(RuntimeError):
    pass


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise BtbbError(f"libbtbb_b200 error {rc}: {lib().btbb_b200_last_error().decode()}")
    return rc


def synth_cfg(n_symbols, stride=10000, n_laps=64, ber=0.0, mix=("DM1", "DM3", "DH1", "FHS"),
              seed=DEFAULT_SEED, first_symbol=0, fixed_lap=0x9E8B33, piconets=False):
    m = 0
    for k in mix:
        m |= 1 << KIND[k]
    return SynthCfg(seed=seed, n_symbols=n_symbols, first_symbol=first_symbol, stride=stride, n_laps=n_laps,
                    ber_q32=min(int(ber * 2 ** 32), 2 ** 32 - 1), packet_mix=m, fixed_lap=fixed_lap, reserved=1 if piconets else 0)


def group_by_lap(hits):
    """btbb_b200_group_by_lap: (order, group_start, laps)."""
    assert hits.dtype == HIT_DTYPE
    n = len(hits)
    order = np.zeros(n, dtype=np.int64)
    gs = np.zeros(n + 1, dtype=np.int64)
    laps = np.zeros(max(n, 1), dtype=np.uint32)
    g = lib().btbb_b200_group_by_lap(hits.ctypes.data, n, order.ctypes.data, gs.ctypes.data, laps.ctypes.data)
    assert g >= 0
    return order, gs[: g + 1].copy(), laps[:g].copy()


def pcap_bredr(hits, dec, meta, reflap=LAP_ANY, refuap=0xFF):
    """File header + records as bytes (btbb_b200_pcap_file_header / _bredr_records)."""
    assert hits.dtype == HIT_DTYPE and dec.dtype == DECODED_DTYPE and meta.dtype == PCAP_META_DTYPE
    assert len(hits) == len(dec) == len(meta)
    L = lib()
    need = L.btbb_b200_pcap_bredr_records(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), reflap, refuap, None, 0)
    buf = np.zeros(24 + need, dtype=np.uint8)
    assert L.btbb_b200_pcap_file_header(buf.ctypes.data, 24) == 24
    got = L.btbb_b200_pcap_bredr_records(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), reflap, refuap,
                                         buf.ctypes.data + 24, need)
    assert got == need
    return buf.tobytes()


OPT_TILE_KERNEL_ONLY, OPT_HOST_BYTE_ROUTE, OPT_HOST_SPLIT_PERMILLE, OPT_PACK_THREADS, OPT_TRACE, OPT_DECODE_WIDE_STAGING, OPT_PACK_STREAMS = 1, 2, 3, 4, 5, 6, 7
MODE_DECODE, MODE_TRY_CLOCKS, MODE_PAYLOAD, MODE_CRC_CHECK, MODE_RAW, MODE_FLAG_RAW_PAYLOAD = 0, 1, 2, 3, 16, 0x100


def decode_smallcall(symbols, length, clkn=0, uap=0, whitened=1, ptype=0, mode=0):
    """btbb_b200_decode_smallcall: the chain for one packet on the host (1 record, 64 in mode 1)."""
    assert symbols.dtype == np.uint8 and symbols.flags.c_contiguous
    out = np.zeros(64 if (mode & 0xff) == MODE_TRY_CLOCKS else 1, dtype=DECODED_DTYPE)
    check(lib().btbb_b200_decode_smallcall(symbols.ctypes.data, length, clkn, uap, whitened, ptype, mode, out.ctypes.data))
    return out


def hop_cfg(address, afh_map=None, aliased=False):
    cfg = HopCfg(address=address & 0xFFFFFFF, afh=1 if afh_map else 0, aliased=1 if aliased else 0)
    if afh_map:
        for i, b in enumerate(bytes(afh_map)[:10]):
            cfg.afh_map[i] = b
    return cfg


def hop_winnow(ctx, cfg, known6, indices, channels, max_candidates=1 << 21):
    """btbb_b200_hop_winnow: (survivors ascending, survivors_after per observation)"""
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    ch = np.ascontiguousarray(channels, dtype=np.uint8)
    cands = np.zeros(max_candidates, dtype=np.uint32)
    after = np.zeros(len(idx), dtype=np.int32)
    n = _i64(0)
    check(lib().btbb_b200_hop_winnow(ctx.h, C.byref(cfg), known6, len(idx), idx.ctypes.data, ch.ctypes.data, cands.ctypes.data,
                                     max_candidates, C.byref(n), after.ctypes.data))
    return cands[: n.value].copy(), after


def pcapng_bredr_blocks(hits, dec, meta, reflap=LAP_ANY, refuap=0xFF):
    """btbb_b200_pcapng_bredr_blocks: the enhanced packet blocks as bytes"""
    L = lib()
    need = L.btbb_b200_pcapng_bredr_blocks(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), reflap, refuap, None, 0)
    buf = np.zeros(max(need, 1), dtype=np.uint8)
    assert L.btbb_b200_pcapng_bredr_blocks(hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), reflap, refuap,
                                           buf.ctypes.data, need) == need
    return buf[:need].tobytes()


def synth_host(cfg):
    buf = np.empty(cfg.n_symbols, dtype=np.uint8)
    check(lib().btbb_b200_synth_host(C.byref(cfg), buf.ctypes.data))
    return buf


def planted(cfg, slot):
    p = Planted()
    check(lib().btbb_b200_synth_planted(C.byref(cfg), slot, C.byref(p)))
    return p


class Context:
    """One CUDA context of the library on one device (btbb_init replacement)."""

    def __init__(self, device=0, max_ac_errors=2):
        self.h = _vp()
        check(lib().btbb_b200_create(device, max_ac_errors, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().btbb_b200_destroy(self.h)
            self.h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def find_ac_host(self, stream, search_length, lap=LAP_ANY, k=2, max_hits=1 << 20):
        """stream: uint8 numpy array holding search_length + 63 symbols."""
        assert stream.dtype == np.uint8 and stream.flags.c_contiguous and len(stream) >= search_length + 63
        hits = np.zeros(max_hits, dtype=HIT_DTYPE)
        n = _i64(0)
        check(lib().btbb_b200_find_ac_host(self.h, stream.ctypes.data, search_length, lap, k,
                                           hits.ctypes.data, max_hits, C.byref(n)))
        return hits[: n.value]

    def find_ac_dev(self, d_ptr, search_length, d_hits_ptr, max_hits, lap=LAP_ANY, k=2, stream=0):
        n = _i64(0)
        rc = lib().btbb_b200_find_ac_dev(self.h, d_ptr, search_length, lap, k, d_hits_ptr, max_hits,
                                         C.byref(n), stream)
        check(rc, allow=(-4,))
        return n.value, rc

    def set_option(self, option, value):
        check(lib().btbb_b200_set_option(self.h, option, value))

    def set_profiling(self, on=True):
        check(lib().btbb_b200_set_profiling(self.h, 1 if on else 0))

    def last_scan_kernel_ms(self):
        ms = C.c_float(0)
        check(lib().btbb_b200_last_scan_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def set_offset_bias(self, bias):
        check(lib().btbb_b200_set_offset_bias(self.h, bias))

    def find_ac_dev_begin(self, d_ptr, search_length, d_hits_ptr, max_hits, lap=LAP_ANY, k=2, stream=0):
        check(lib().btbb_b200_find_ac_dev_begin(self.h, d_ptr, search_length, lap, k, d_hits_ptr, max_hits, stream))

    def find_ac_dev_end(self):
        n = _i64(0)
        rc = lib().btbb_b200_find_ac_dev_end(self.h, C.byref(n))
        check(rc, allow=(-4,))
        return n.value, rc

    def find_ac_packed_dev(self, d_words_ptr, search_length, d_hits_ptr, max_hits, lap=LAP_ANY, k=2, stream=0):
        n = _i64(0)
        rc = lib().btbb_b200_find_ac_packed_dev(self.h, d_words_ptr, search_length, lap, k, d_hits_ptr, max_hits,
                                                C.byref(n), stream)
        check(rc, allow=(-4,))
        return n.value, rc

    def uap_sieve_host(self, stream, pkts, group_start, states):
        """btbb_b200_uap_sieve_host: returns (updated states, rv per packet)."""
        assert stream.dtype == np.uint8 and pkts.dtype == PKTIN_DTYPE and states.dtype == SIEVE_DTYPE
        gs = np.ascontiguousarray(group_start, dtype=np.int64)
        assert len(gs) == len(states) + 1 and gs[-1] <= len(pkts)
        st = states.copy()
        rv = np.zeros(len(pkts), dtype=np.int8)
        check(lib().btbb_b200_uap_sieve_host(self.h, stream.ctypes.data, len(stream), pkts.ctypes.data, len(pkts),
                                             gs.ctypes.data, len(st), st.ctypes.data, rv.ctypes.data))
        return st, rv

    def decode_host(self, stream, pkts, mode=0):
        assert stream.dtype == np.uint8 and pkts.dtype == PKTIN_DTYPE
        out = np.zeros(len(pkts) * (64 if (mode & 0xff) == MODE_TRY_CLOCKS else 1), dtype=DECODED_DTYPE)
        check(lib().btbb_b200_decode_host(self.h, stream.ctypes.data, len(stream), pkts.ctypes.data, len(pkts),
                                          mode, out.ctypes.data))
        return out
