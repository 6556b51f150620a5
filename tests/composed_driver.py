"""Drives ONE libbtbb build through the public btbb.h API (helper of tests/test_composed.py, run as
a subprocess once per library): survey mode and piconet-following btbb_process_packet on a
synthetic capture, then the pcap / pcapng writers.  Prints one JSON object.

    python tests/composed_driver.py <path to libbtbb.so>
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import util  # noqa: E402
from util import B  # noqa: E402

def pcapng_normalised(data):
    """upstream's btbb_pcapng_append_packet assembles the block in an uninitialised stack struct
    (pcapng-bt.c:244-246), so the <= 3 pad bytes behind each packet's data are whatever the stack
    held: zero them before comparing two files"""
    b = bytearray(data)
    at = 0
    while at + 12 <= len(b):
        btype = int.from_bytes(b[at:at + 4], "little")
        blen = int.from_bytes(b[at + 4:at + 8], "little")
        if blen < 12 or at + blen > len(b):
            break
        if btype == 6:      # enhanced packet block
            cap = int.from_bytes(b[at + 20:at + 24], "little")
            for i in range(at + 28 + cap, at + blen - 8):
                b[i] = 0
        at += blen
    return bytes(b)


LAP_OFFSET = 12      # struct btbb_packet: refcount, flags, channel, UAP, NAP, LAP (bluetooth_packet.h:52-66)


def main(path):
    L = C.CDLL(path)
    L.btbb_packet_new.restype = C.c_void_p
    L.btbb_packet_unref.argtypes = [C.c_void_p]
    L.btbb_packet_set_flag.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.btbb_packet_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint8, C.c_uint32]
    L.btbb_packet_set_uap.argtypes = [C.c_void_p, C.c_uint8]
    L.btbb_decode_header.argtypes = [C.c_void_p]
    L.btbb_decode_payload.argtypes = [C.c_void_p]
    L.btbb_process_packet.argtypes = [C.c_void_p, C.c_void_p]
    L.btbb_piconet_new.restype = C.c_void_p
    L.btbb_init_piconet.argtypes = [C.c_void_p, C.c_uint32]
    L.btbb_next_survey_result.restype = C.c_void_p
    L.btbb_piconet_get_lap.restype = C.c_uint32
    L.btbb_piconet_get_uap.restype = C.c_uint8
    L.btbb_piconet_get_afh_map.restype = C.POINTER(C.c_uint8)
    for f in ("btbb_piconet_get_lap", "btbb_piconet_get_uap", "btbb_piconet_get_clk_offset", "btbb_piconet_get_afh_map"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.btbb_piconet_get_flag.argtypes = [C.c_void_p, C.c_int]
    L.btbb_pcap_create_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.btbb_pcap_append_packet.argtypes = [C.c_void_p, C.c_uint64, C.c_int8, C.c_int8, C.c_uint32, C.c_uint8, C.c_void_p]
    L.btbb_pcap_close.argtypes = [C.c_void_p]
    L.btbb_pcapng_create_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.btbb_pcapng_append_packet.argtypes = [C.c_void_p, C.c_uint64, C.c_int8, C.c_int8, C.c_uint32, C.c_uint8, C.c_void_p]
    L.btbb_pcapng_close.argtypes = [C.c_void_p]

    stream, pkts, gs, laps, truth = util.sieve_case(n_slots=260, stride=4000, n_laps=5, ber=0.002, seed=77)
    lap_of = np.zeros(len(pkts), dtype=np.uint32)
    for g, lap in enumerate(laps):
        lap_of[gs[g]:gs[g + 1]] = lap
    order = np.argsort(pkts["offset"], kind="stable")

    def make_packet(i):
        p = L.btbb_packet_new()
        C.c_uint32.from_address(p + LAP_OFFSET).value = int(lap_of[i])
        L.btbb_packet_set_flag(p, 0, 1)      # BTBB_WHITENED, as init_packet sets it (:201-208)
        q = pkts[i]
        L.btbb_packet_set_data(p, stream.ctypes.data + int(q["offset"]), int(q["length"]), int(q["reserved"]) & 0xff, int(q["clkn"]) << 1)
        return p

    out = {}
    # library chatter ("UAP = .. found", btbb_decode's report) goes to /dev/null
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        # ---- survey mode (bluetooth_piconet.c:851-858) ----
        import time
        L.btbb_init_survey()
        t0 = time.perf_counter()
        for i in order:
            p = make_packet(i)
            L.btbb_process_packet(p, None)
            L.btbb_packet_unref(p)
        out["survey_us_per_packet"] = 1e6 * (time.perf_counter() - t0) / max(len(order), 1)
        survey = []
        while True:
            pn = L.btbb_next_survey_result()
            if not pn:
                break
            afh = bytes(L.btbb_piconet_get_afh_map(pn)[:10]).hex()
            survey.append([L.btbb_piconet_get_lap(pn), L.btbb_piconet_get_uap(pn), L.btbb_piconet_get_flag(pn, 2),
                           L.btbb_piconet_get_clk_offset(pn), afh])
        out["survey"] = sorted(survey)
        # ---- pcap / pcapng of packets decoded with their true clock / UAP ----
        for kind in ("pcap", "pcapng"):
            path = f"/tmp/composed_{os.getpid()}.{kind}".encode()
            h = C.c_void_p(None)
            if kind == "pcap":
                assert L.btbb_pcap_create_file(path, C.byref(h)) == 0
            else:
                assert L.btbb_pcapng_create_file(path, b"composed driver", C.byref(h)) == 0
            for n, i in enumerate(order[:120]):
                p = make_packet(i)
                uap, clk0 = truth[int(lap_of[i])]
                slot = int(pkts[i]["clkn"])
                L.btbb_packet_set_uap(p, uap)
                L.btbb_packet_set_data(p, stream.ctypes.data + int(pkts[i]["offset"]), int(pkts[i]["length"]),
                                       int(pkts[i]["reserved"]) & 0xff, ((clk0 + slot) & 63) << 1)
                L.btbb_packet_set_flag(p, 4, 1)      # BTBB_CLK6_VALID
                if L.btbb_decode_header(p):
                    L.btbb_decode_payload(p)
                ap = L.btbb_pcap_append_packet if kind == "pcap" else L.btbb_pcapng_append_packet
                ap(h, 1_000_000 * n, -40 - n % 30, -90, 0xFFFFFFFF if n % 2 else int(lap_of[i]), 0xFF if n % 2 else uap, p)
                L.btbb_packet_unref(p)
            (L.btbb_pcap_close if kind == "pcap" else L.btbb_pcapng_close)(h)
            data = open(path.decode(), "rb").read()
            os.remove(path.decode())
            if kind == "pcapng":
                data = pcapng_normalised(data)
            out[kind] = [len(data), hashlib.sha256(data).hexdigest()]
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
