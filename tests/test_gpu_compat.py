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


C_COMPUTE_CALLER = r"""
/* an existing libbtbb caller, plain C99, linked with -lbtbb: the batch C ABI and the classic calls
 * doing real work on the GPU -- prints what it found for the test to compare with the oracle */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "btbb.h"
#include "btbb_b200.h"

int main(void)
{
	btbb_b200_ctx *ctx = NULL;
	btbb_b200_synth_cfg cfg;
	btbb_b200_hit *hits;
	btbb_packet *pkt = NULL;
	int64_t n = 3000000, cap = 4096, got = 0, i;
	unsigned long long sum = 0;
	uint8_t *s;
	int off, rv;
	if (btbb_b200_create(0, 2, &ctx) != BTBB_B200_OK) { fprintf(stderr, "%s\n", btbb_b200_last_error()); return 1; }
	memset(&cfg, 0, sizeof(cfg));
	cfg.seed = 4711; cfg.n_symbols = n + 63; cfg.stride = 5000; cfg.n_laps = 64;
	cfg.ber_q32 = 4294967;      /* 0.1 % */
	cfg.packet_mix = (1u << BTBB_B200_KIND_DM1) | (1u << BTBB_B200_KIND_FHS) | (1u << BTBB_B200_KIND_DH1);
	cfg.fixed_lap = 0x9e8b33;
	s = malloc((size_t)n + 63);
	hits = malloc((size_t)cap * sizeof(*hits));
	if (!s || !hits || btbb_b200_synth_host(&cfg, s) != BTBB_B200_OK) return 2;
	if (btbb_b200_find_ac_host(ctx, (const char *)s, n, BTBB_B200_LAP_ANY, 2, hits, cap, &got) != BTBB_B200_OK) return 3;
	for (i = 0; i < got; i++)
		sum = sum * 1000003ull + (unsigned long long)hits[i].offset * 31ull + hits[i].lap * 7ull + hits[i].ac_errors;
	printf("%lld %llu\n", (long long)got, sum);
	/* the classic surface: a search long enough for the kernels, then the per-packet calls */
	if (btbb_init(2) != 0) return 4;
	off = btbb_find_ac((char *)s, 100000, LAP_ANY, 2, &pkt);
	printf("%d %06x %d\n", off, pkt ? btbb_packet_get_lap(pkt) : 0u, pkt ? btbb_packet_get_ac_errors(pkt) : -1);
	if (off >= 0) {
		btbb_b200_planted pl;
		btbb_b200_synth_planted(&cfg, off / cfg.stride, &pl);
		btbb_packet_set_data(pkt, (char *)s + off, 3125, 0, (uint32_t)pl.clk6 << 1);
		btbb_packet_set_uap(pkt, pl.uap);
		btbb_packet_set_flag(pkt, BTBB_CLK6_VALID, 1);
		rv = btbb_decode_header(pkt) ? btbb_decode_payload(pkt) : -1;
		printf("%d %d %d\n", rv, btbb_packet_get_type(pkt), btbb_packet_get_payload_length(pkt));
		btbb_packet_unref(pkt);
	}
	btbb_b200_destroy(ctx);
	free(s); free(hits);
	return 0;
}
"""


def test_c99_program_computes_through_the_library(product_lib, orc, tmp_path):
    """tests/test_abi.py builds a C caller that only creates a context; this one scans, searches and
    decodes: a C99 program linked against libbtbb.so.1 by SONAME, its output against the oracle."""
    import os
    import subprocess
    src = tmp_path / "compute.c"
    src.write_text(C_COMPUTE_CALLER)
    exe = tmp_path / "compute"
    libdir = os.path.dirname(B.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(util.ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:libbtbb.so.1", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.strip().splitlines() if not l.startswith("Packet decoded") and not l.startswith("  ")]
    n = 3000000
    cfg = B.synth_cfg(n + 63, stride=5000, ber=4294967 / 2 ** 32, seed=4711, mix=("DM1", "FHS", "DH1"))
    s = B.synth_host(cfg)
    orc.orc_init(2)
    want = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    acc = 0
    for h in want:
        acc = (acc * 1000003 + int(h["offset"]) * 31 + int(h["lap"]) * 7 + int(h["ac_errors"])) % (1 << 64)
    assert int(lines[0][0]) == len(want) > 500 and int(lines[0][1]) == acc
    first = want[0]
    assert int(lines[1][0]) == int(first["offset"]) and int(lines[1][1], 16) == int(first["lap"]) and int(lines[1][2]) == int(first["ac_errors"])
    p = B.planted(cfg, int(first["offset"]) // 5000)
    d = util.decode_one(orc, "orc", s, int(first["offset"]), 3125, p.clk6, p.uap)
    assert [int(x) for x in lines[2]] == [int(d["rv"]) if d["header_ok"] else -1, int(d["type"]), int(d["payload_length"])]
