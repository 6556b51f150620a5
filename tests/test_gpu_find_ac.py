"""Parity of the CUDA access-code correlator with the oracle / golden fixtures, through the
C ABI (btbb_b200_find_ac_host / _dev).  Bit-exact: every hit record, in order."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import util
from util import B

sys.path.insert(0, util.GOLDEN)
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rand_stream():
    return make_golden.find_ac_stream(1234, 1 << 20)


@pytest.mark.parametrize("k_init", [0, 1, 2, 3, 4])
def test_golden_fixture_cases(product_lib, rand_stream, k_init):
    """Outputs recorded from the unmodified reference for btbb_init(k_init)."""
    cases = json.load(open(os.path.join(util.GOLDEN, "find_ac.json")))[k_init]["cases"]
    cfg, synth = make_golden.synth_stream(0.005)
    with B.Context(0, k_init) as ctx:
        assert product_lib.btbb_b200_table_errors(ctx.h) == k_init
        for c in cases:
            s = rand_stream if c["stream"] == "rand1234" else synth
            h = ctx.find_ac_host(s, c["n"], c["lap"], c["k"])
            assert len(h) == c["count"], c
            assert util.digest(h) == c["sha256"], c


def test_oracle_promiscuous_and_known(gpu_ctx2, orc):
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(2024)
    n = 3_000_017
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    util.plant_syncwords(s, rng, 400, 3)
    util.plant_syncwords(s, rng, 100, 8, laps=[0x123456])
    for lap, k in [(B.LAP_ANY, 0), (B.LAP_ANY, 1), (B.LAP_ANY, 2), (B.LAP_ANY, 5), (B.LAP_ANY, -1),
                   (0x123456, 0), (0x123456, 3), (0x123456, 8), (0x123456, 17), (0x123456, -1)]:
        want = util.find_all(orc, "orc", s, n, lap, k)
        got = gpu_ctx2.find_ac_host(s, n, lap, k)
        assert got.tobytes() == want.tobytes(), (hex(lap), k, len(got), len(want))


def test_dense_hits_known_lap(gpu_ctx2, orc):
    """k = 24..64: a large share of all positions hit; exercises the append + ordering path."""
    rng = np.random.default_rng(5)
    n = 200_000
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    for k in (24, 32, 64):
        want = util.find_all(orc, "orc", s, n, 0x9E8B33, k)
        got = gpu_ctx2.find_ac_host(s, n, 0x9E8B33, k, max_hits=n)
        assert len(want) > n // 1000 and got.tobytes() == want.tobytes()
    assert len(gpu_ctx2.find_ac_host(s, n, 0x9E8B33, 64, max_hits=n)) == n


def _barker_dense_streams(rng, n):
    """Streams on which far more than 1/8 of all positions pass the Barker-tail filter, up to
    nearly all of them: they overflow the bulk kernel's per-warp candidate queue and force its
    in-place fallback."""
    out = {}
    a = np.array([(0x27 >> i) & 1 for i in range(7)], dtype=np.uint8)         # tail A, period 7
    out["tailA_period7"] = np.resize(a, n + 63)
    out["ones"] = np.ones(n + 63, dtype=np.uint8)
    out["zeros"] = np.zeros(n + 63, dtype=np.uint8)
    s = np.resize(a, n + 63).copy()                                           # tail A with 3 % flips
    s ^= (rng.random(n + 63) < 0.03).astype(np.uint8)
    out["tailA_noisy"] = s
    words = []                                                                # back-to-back sync words
    laps = rng.integers(0, 1 << 24, (n + 63) // 64 + 1)
    for lap in laps:
        w = int(B.lib().btbb_gen_syncword(int(lap))) & 0xFFFFFFFFFFFFFFFF
        words.append([(w >> i) & 1 for i in range(64)])
    s = np.array(words, dtype=np.uint8).reshape(-1)[: n + 63].copy()
    s ^= (rng.random(n + 63) < 0.01).astype(np.uint8)
    out["syncwords_back_to_back"] = s
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)                            # half dense, half noise
    s[: n // 2] = np.resize(a, n // 2)
    out["half_dense"] = s
    return out


def test_barker_dense_streams(gpu_ctx2, orc, product_lib):
    assert orc.orc_init(2) == 0
    product_lib.btbb_gen_syncword.restype = C.c_uint64
    rng = np.random.default_rng(77)
    n = 150_000
    for name, s in _barker_dense_streams(rng, n).items():
        for k in (0, 2):
            want = util.find_all(orc, "orc", s, n, B.LAP_ANY, k)
            got = gpu_ctx2.find_ac_host(s, n, B.LAP_ANY, k, max_hits=n)
            assert got.tobytes() == want.tobytes(), (name, k, len(got), len(want))


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 63, 64, 65, 8191, 8192, 8193, 16384 + 5])
def test_edge_lengths_and_seams(gpu_ctx2, orc, n):
    """Tiny and tile-boundary lengths; hits planted at the first/last position and on seams."""
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(n + 1)
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    sw = orc.orc_gen_syncword(0x9E8B33)
    bits = [(sw >> i) & 1 for i in range(64)]
    for p in (0, n - 1, 8191, 8192 - 40, 8160, 31, 32):
        if 0 <= p < n:
            s[p:p + 64] = bits
    for lap, k in [(B.LAP_ANY, 2), (0x9E8B33, 1)]:
        want = util.find_all(orc, "orc", s, n, lap, k) if n > 0 else np.zeros(0, dtype=B.HIT_DTYPE)
        got = gpu_ctx2.find_ac_host(s, n, lap, k)
        assert got.tobytes() == want.tobytes(), (n, hex(lap), len(got), len(want))
        if n > 0:
            assert len(want) >= 1


def test_device_api_alignment_and_overflow(gpu_ctx2, orc, product_lib):
    import torch
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(11)
    n = 500_000
    s = rng.integers(0, 2, n + 63 + 16, dtype=np.uint8)
    util.plant_syncwords(s, rng, 300, 2)
    d = torch.from_numpy(s).cuda()
    d_hits = torch.zeros(4096 * 16, dtype=torch.uint8, device="cuda")
    for head in (0, 1, 7, 15):           # stream pointer not 16-byte aligned
        want = util.find_all(orc, "orc", s[head:], n, B.LAP_ANY, 2)
        cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr() + head, n, d_hits.data_ptr(), 4096,
                                       stream=torch.cuda.current_stream().cuda_stream)
        assert rc == 0 and cnt == len(want)
        got = d_hits.cpu().numpy()[: cnt * 16].view(B.HIT_DTYPE)
        assert got.tobytes() == want.tobytes()
    want = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), 10)   # buffer too small
    assert rc == -4 and cnt == len(want)
    got = d_hits.cpu().numpy()[: 160].view(B.HIT_DTYPE)
    assert np.all(np.diff(got["offset"]) > 0) and set(got["offset"].tolist()) <= set(want["offset"].tolist())


def test_synth_device_equals_host(gpu_ctx2, product_lib):
    import torch
    for first, n, ber in ((0, 1_000_003, 0.0), (123_457, 400_000, 0.01)):
        cfg = B.synth_cfg(n, stride=4096, ber=ber, first_symbol=first, mix=tuple(B.KIND))
        host = B.synth_host(cfg)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        B.check(product_lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), host)


def test_full_size_10gbit_properties(product_lib):
    """BASELINE configs[1] at full size (10^10 symbols): size-independent properties.
    (1) every planted access code (BER 0) is reported at its exact offset with its LAP;
    (2) scanning two halves with a 63-symbol seam gives the same list as one scan;
    (3) the extra hits are as rare as the false-positive rate of the code predicts."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    n = 10**10 if free > 13 * 2**30 else 2 * 10**9
    stride = 10000
    cfg = B.synth_cfg(n + 63, stride=stride, ber=0.0, mix=("ID", "DM1", "DM3", "DH1", "FHS"))
    d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
    B.check(product_lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0))
    torch.cuda.synchronize()
    cap = n // stride + 100_000
    d_hits = torch.zeros(cap * 16, dtype=torch.uint8, device="cuda")
    with B.Context(0, 2) as ctx:
        cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap)
        assert rc == 0
        hits = d_hits.cpu().numpy()[: cnt * 16].view(B.HIT_DTYPE).copy()
        assert np.all(np.diff(hits["offset"]) > 0)
        # (1) planted ground truth, checked on a sample of slots spread over the stream
        offs = hits["offset"]
        nslots = n // stride
        for slot in np.linspace(0, nslots - 2, 4000).astype(np.int64):
            p = B.planted(cfg, int(slot))
            i = np.searchsorted(offs, p.offset)
            assert i < cnt and offs[i] == p.offset and hits["lap"][i] == p.lap and hits["ac_errors"][i] == 0
        extra = cnt - nslots
        assert 0 <= extra < max(2000, int(n * 5e-8))          # ~1.2e-8 false hits per symbol at k=2
        # (2) seam invariance
        half = n // 2 + 12345
        c1, _ = ctx.find_ac_dev(d.data_ptr(), half, d_hits.data_ptr(), cap)
        h1 = d_hits.cpu().numpy()[: c1 * 16].view(B.HIT_DTYPE).copy()
        c2, _ = ctx.find_ac_dev(d.data_ptr() + half, n - half, d_hits.data_ptr(), cap)
        h2 = d_hits.cpu().numpy()[: c2 * 16].view(B.HIT_DTYPE).copy()
        h2["offset"] += half
        assert np.concatenate([h1, h2]).tobytes() == hits.tobytes()


def _pack_words(s):
    """format B: symbol i -> bit (i & 31) of word i >> 5 (bluetooth_packet.c:235-242 bit order)"""
    pad = (-len(s)) % 32
    b = np.packbits(np.concatenate([s, np.zeros(pad, dtype=np.uint8)]), bitorder="little")
    return b.view("<u4").copy()


@pytest.mark.parametrize("n", [1, 33, 4095, 4096, 4097, 4160, 8192 + 7, 300_001, 2_000_003])
def test_packed_ingest_matches_oracle(gpu_ctx2, orc, n):
    """btbb_b200_find_ac_packed_dev == oracle on the same symbols, all bulk / tail / fallback routes."""
    import torch
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(77 + n)
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    if n > 100:
        util.plant_syncwords(s, rng, max(1, n // 5000), 2)
        util.plant_syncwords(s, rng, max(1, n // 20000), 4, laps=[0x123456])
    sw = orc.orc_gen_syncword(0x123456)
    for p in (0, n - 1, 4095, 4096, 4064):          # first / last position and strip seams
        if 0 <= p < n:
            s[p:p + 64] = [(sw >> i) & 1 for i in range(64)]
    d_words = torch.from_numpy(_pack_words(s).view(np.int32)).cuda()
    cap = n + 16
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    for lap, k in [(B.LAP_ANY, 2), (B.LAP_ANY, 1), (B.LAP_ANY, 4), (0x123456, 2), (0x123456, 5), (0x123456, 20)]:
        want = util.find_all(orc, "orc", s, n, lap, k)
        cnt, rc = gpu_ctx2.find_ac_packed_dev(d_words.data_ptr(), n, d_hits.data_ptr(), cap, lap=lap, k=k)
        assert rc == 0 and cnt == len(want), (n, hex(lap), k, cnt, len(want))
        got = d_hits[:cnt].cpu().numpy().tobytes()
        assert got == want.tobytes(), (n, hex(lap), k)


def test_begin_end_halves_and_offset_bias(gpu_ctx2, orc):
    """btbb_b200_find_ac_dev == _begin + _end; the offset bias shifts every reported offset (also on the
    generic ordering path with a bias that makes the offsets straddle a power of two)."""
    import torch
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(4242)
    n = 500_003
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    util.plant_syncwords(s, rng, 120, 2)
    want = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    d = torch.from_numpy(s).cuda()
    cap = 4096
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    for lap in (B.LAP_ANY, int(want[0]["lap"])):
        ref = want if lap == B.LAP_ANY else util.find_all(orc, "orc", s, n, lap, 2)
        gpu_ctx2.find_ac_dev_begin(d.data_ptr(), n, d_hits.data_ptr(), cap, lap=lap, k=2)
        cnt, rc = gpu_ctx2.find_ac_dev_end()
        assert rc == 0 and d_hits[:cnt].cpu().numpy().tobytes() == ref.tobytes()
    with pytest.raises(B.BtbbError):
        gpu_ctx2.find_ac_dev_end()                      # nothing pending
    try:
        gpu_ctx2.set_offset_bias(10**11)
        cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=2)
        got = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE).copy()
        assert rc == 0 and cnt == len(want)
        assert np.array_equal(got["offset"], want["offset"] + 10**11)
        got["offset"] -= 10**11
        assert got.tobytes() == want.tobytes()
        h = gpu_ctx2.find_ac_host(s, n, B.LAP_ANY, 2)   # the host entry point is not biased
        assert h.tobytes() == want.tobytes()
        # known LAP takes the radix-sort path: digits are taken from offset - bias (a bias that
        # carries the offsets across 2^18 / 2^27 used to wrap the low digits)
        lap = int(want[0]["lap"])
        ref = util.find_all(orc, "orc", s, n, lap, 4)
        for bias in ((1 << 18) - 1000, (1 << 27) - n // 2, 10**11 + 7, -(n // 3)):
            gpu_ctx2.set_offset_bias(bias)
            cnt, rc = gpu_ctx2.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, lap=lap, k=4)
            got = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE).copy()
            assert rc == 0 and np.array_equal(got["offset"], ref["offset"] + bias), bias
    finally:
        gpu_ctx2.set_offset_bias(0)


@pytest.mark.parametrize("k_init", [3, 4])
def test_packed_and_byte_ingest_with_larger_tables(product_lib, k_init):
    """Tables for 3 / 4 errors take the bulk kernel's global-memory map variants, for both input
    formats; outputs recorded from the reference in a subprocess-free way: the oracle with the
    same table size."""
    import subprocess, sys, torch
    n = 300_001
    rng = np.random.default_rng(500 + k_init)
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    util.plant_syncwords(s, rng, 60, k_init)
    # the oracle's table is built once per process (like the reference's): ask a fresh interpreter
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import util; from util import B; "
            "O = util.oracle(); assert O.orc_init(%d) == 0; s = np.load(sys.argv[1]); "
            "h = util.find_all(O, 'orc', s, %d, B.LAP_ANY, %d); np.save(sys.argv[2], h)") % (
                util.ROOT, os.path.join(util.ROOT, "tests"), k_init, n, k_init)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "s.npy"), s)
        subprocess.run([sys.executable, "-c", code, os.path.join(td, "s.npy"), os.path.join(td, "h.npy")], check=True)
        want = np.load(os.path.join(td, "h.npy"))
    assert len(want) >= 40          # (errors planted in the Barker tail can put a sync word out of reach)
    d_words = torch.from_numpy(_pack_words(s).view(np.int32)).cuda()
    d_bytes = torch.from_numpy(s).cuda()
    cap = 1 << 16
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    with B.Context(0, k_init) as ctx:
        cnt, rc = ctx.find_ac_packed_dev(d_words.data_ptr(), n, d_hits.data_ptr(), cap, k=k_init)
        assert rc == 0 and d_hits[:cnt].cpu().numpy().tobytes() == want.tobytes(), ("packed", cnt, len(want))
        cnt, rc = ctx.find_ac_dev(d_bytes.data_ptr(), n, d_hits.data_ptr(), cap, k=k_init)
        assert rc == 0 and d_hits[:cnt].cpu().numpy().tobytes() == want.tobytes(), ("bytes", cnt, len(want))


def test_host_entry_point_large_buffer_packs(gpu_ctx2, orc):
    """>= 4 Mi symbols: find_ac_host packs on the host before the copy; the byte-format copy
    (BTBB_B200_OPT_HOST_BYTE_ROUTE) and the oracle must give the same records, pageable memory included."""
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(99)
    n = 6_000_011
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    util.plant_syncwords(s, rng, 900, 2)
    util.plant_syncwords(s, rng, 200, 3, laps=[0x9E8B33])
    for lap, k in [(B.LAP_ANY, 2), (0x9E8B33, 3), (B.LAP_ANY, 3)]:
        want = util.find_all(orc, "orc", s, n, lap, k)
        got = gpu_ctx2.find_ac_host(s, n, lap, k)
        assert got.tobytes() == want.tobytes(), (hex(lap), k, len(got), len(want))
        gpu_ctx2.set_option(B.OPT_HOST_BYTE_ROUTE, 1)
        got2 = gpu_ctx2.find_ac_host(s, n, lap, k)
        gpu_ctx2.set_option(B.OPT_HOST_BYTE_ROUTE, 0)
        assert got2.tobytes() == want.tobytes(), ("bytes", hex(lap), k)
        # head as bytes over DMA, rest packed (what a pinned buffer gets), forced split points
        for frac in (370, 1, 950):
            gpu_ctx2.set_option(B.OPT_HOST_SPLIT_PERMILLE, frac)
            got3 = gpu_ctx2.find_ac_host(s, n, lap, k)
            gpu_ctx2.set_option(B.OPT_HOST_SPLIT_PERMILLE, 0)
            assert got3.tobytes() == want.tobytes(), ("split", frac, hex(lap), k)
    import torch
    pinned = torch.from_numpy(s).pin_memory()
    want = util.find_all(orc, "orc", s, n, B.LAP_ANY, 2)
    got4 = gpu_ctx2.find_ac_host(pinned.numpy(), n, B.LAP_ANY, 2)
    assert got4.tobytes() == want.tobytes()
    # a hit buffer that is too small: EOVERFLOW, the full count, and max_hits genuine records in
    # ascending order (which ones is not specified) on every route
    import ctypes as C
    few = np.zeros(50, dtype=B.HIT_DTYPE)
    cnt = C.c_int64(0)
    for env in (0, 500):
        gpu_ctx2.set_option(B.OPT_HOST_SPLIT_PERMILLE, env)
        rc = B.lib().btbb_b200_find_ac_host(gpu_ctx2.h, s.ctypes.data, n, B.LAP_ANY, 2, few.ctypes.data, 50, C.byref(cnt))
        gpu_ctx2.set_option(B.OPT_HOST_SPLIT_PERMILLE, 0)
        assert rc == -4 and cnt.value == len(want), (env, rc, cnt.value)
        wanted = {r.tobytes() for r in want}
        assert all(r.tobytes() in wanted for r in few) and (np.diff(few["offset"]) > 0).all(), env


def test_two_pending_scans_and_tile_kernel_option(gpu_ctx2, orc, product_lib):
    """Two scans may be pending on one context (_end completes the oldest); a third _begin is refused;
    the tile-kernel-only option gives the same list as the bulk kernels."""
    import ctypes as C
    import torch
    assert orc.orc_init(2) == 0
    rng = np.random.default_rng(321)
    n = 3_000_017
    s = rng.integers(0, 2, n + 63, dtype=np.uint8)
    util.plant_syncwords(s, rng, 700, 2)
    d = torch.from_numpy(s).cuda()
    cap = 1 << 16
    ha = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    hb = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    na, nb = n, 1_234_567
    want_a = util.find_all(orc, "orc", s, na, B.LAP_ANY, 2)
    want_b = util.find_all(orc, "orc", s, nb, B.LAP_ANY, 1)
    ctx = gpu_ctx2
    for _ in range(3):
        ctx.find_ac_dev_begin(d.data_ptr(), na, ha.data_ptr(), cap, k=2)
        ctx.find_ac_dev_begin(d.data_ptr(), nb, hb.data_ptr(), cap, k=1)
        assert product_lib.btbb_b200_find_ac_dev_begin(ctx.h, d.data_ptr(), nb, B.LAP_ANY, 1, hb.data_ptr(), cap, 0) == -1
        ca, rc = ctx.find_ac_dev_end()
        assert rc == 0 and ha[:ca].cpu().numpy().tobytes() == want_a.tobytes()
        cb, rc = ctx.find_ac_dev_end()
        assert rc == 0 and hb[:cb].cpu().numpy().tobytes() == want_b.tobytes()
    n_end = C.c_int64(0)
    assert product_lib.btbb_b200_find_ac_dev_end(ctx.h, C.byref(n_end)) == -1      # nothing pending
    # a known-LAP scan (generic ordering path) queued behind a promiscuous one
    lap = int(want_a[0]["lap"])
    want_k = util.find_all(orc, "orc", s, na, lap, 3)
    ctx.find_ac_dev_begin(d.data_ptr(), na, ha.data_ptr(), cap, k=2)
    ctx.find_ac_dev_begin(d.data_ptr(), na, hb.data_ptr(), cap, lap=lap, k=3)
    ca, rc = ctx.find_ac_dev_end()
    cb, rc2 = ctx.find_ac_dev_end()
    assert rc == 0 and rc2 == 0 and ha[:ca].cpu().numpy().tobytes() == want_a.tobytes() and hb[:cb].cpu().numpy().tobytes() == want_k.tobytes()
    ctx.set_option(B.OPT_TILE_KERNEL_ONLY, 1)
    ca, rc = ctx.find_ac_dev(d.data_ptr(), na, ha.data_ptr(), cap, k=2)
    ctx.set_option(B.OPT_TILE_KERNEL_ONLY, 0)
    assert rc == 0 and ha[:ca].cpu().numpy().tobytes() == want_a.tobytes()
