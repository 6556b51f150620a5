"""Parity of the CUDA per-packet chain (unfec13/unwhiten/HEC, unfec23/CRC, all type
decoders) with the oracle and the golden fixtures, through btbb_b200_decode_host."""
import json
import os
import sys

import numpy as np
import pytest

import util
from util import B

sys.path.insert(0, util.GOLDEN)
import make_golden  # noqa: E402
from test_oracle_vs_golden import noise_type_records  # noqa: E402

pytestmark = pytest.mark.gpu


def pkts_for(planted, s, length=None, clk_delta=0, uap_delta=0, whitened=1):
    a = np.zeros(len(planted), dtype=B.PKTIN_DTYPE)
    for i, p in enumerate(planted):
        a[i]["offset"] = p.offset
        a[i]["length"] = min(3125 if length is None else length, len(s) - p.offset)
        a[i]["clkn"] = (p.clk6 + clk_delta) & 63
        a[i]["uap"] = (p.uap + uap_delta) & 255
        a[i]["whitened"] = whitened
    return a


def test_golden_decode_fixture(gpu_ctx2):
    g = json.load(open(os.path.join(util.GOLDEN, "decode.json")))
    for gs in g["streams"]:
        cfg, s = make_golden.synth_stream(gs["ber"])
        pl = util.planted_list(cfg)
        recs = gpu_ctx2.decode_host(s, pkts_for(pl, s), mode=0)
        assert recs[:4].tobytes().hex() == gs["decode_head"]
        assert util.digest(recs) == gs["decode_sha256"]
        tc = gpu_ctx2.decode_host(s, pkts_for(pl[:40], s), mode=1)
        assert util.digest(tc) == gs["try_clock_sha256"]
        odd = []
        for p in pl[:60]:
            for L in (100, 121, 122, 130, 137, 200, 361, 362, 500):
                odd.append(pkts_for([p], s, length=L)[0])
            odd.append(pkts_for([p], s, clk_delta=1)[0])
            odd.append(pkts_for([p], s, uap_delta=1)[0])
            odd.append(pkts_for([p], s, whitened=0)[0])
        got = gpu_ctx2.decode_host(s, np.array(odd, dtype=B.PKTIN_DTYPE), mode=0)
        assert util.digest(got) == gs["odd_sha256"]


def test_oracle_decode_and_try_clocks(gpu_ctx2, orc):
    cfg = B.synth_cfg(2_000_000, stride=3400, ber=0.006, seed=31337, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)
    got = gpu_ctx2.decode_host(s, pkts_for(pl, s), mode=0)
    want = np.array([util.decode_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), p.clk6, p.uap) for p in pl])
    bad = [i for i in range(len(pl)) if got[i].tobytes() != want[i].tobytes()]
    assert bad == [], (bad[:5], got[bad[0]], want[bad[0]])
    sub = pl[:120]
    got = gpu_ctx2.decode_host(s, pkts_for(sub, s), mode=1)
    want = np.array([util.try_clock_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), c)
                     for p in sub for c in range(64)])
    bad = [i for i in range(len(want)) if got[i].tobytes() != want[i].tobytes()]
    assert bad == [], (bad[:5], got[bad[0]], want[bad[0]])


def test_all_packet_types_on_noise(gpu_ctx2):
    """EV3/EV4/EV5/HV2/HV3/DV/AUX1/NULL/POLL included (fixture from the reference)."""
    g = json.load(open(os.path.join(util.GOLDEN, "noise_types.json")))
    rng = np.random.default_rng(g["seed"])
    streams, pk1, pk0, order = [], [], [], []
    for i in range(300):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        sym[68:122] = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        n = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        clk = int(rng.integers(0, 64))
        streams.append(sym)
        a = np.zeros(1, dtype=B.PKTIN_DTYPE)[0]
        a["offset"], a["length"], a["clkn"], a["uap"], a["whitened"] = i * 3125, n, clk, 0, 1
        pk0.append(a.copy())
    s = np.concatenate(streams)
    pk = np.array(pk0, dtype=B.PKTIN_DTYPE)
    m1 = gpu_ctx2.decode_host(s, pk, mode=1).reshape(300, 64)
    m0 = gpu_ctx2.decode_host(s, pk, mode=0)
    recs = []
    for i in range(300):
        recs.extend(m1[i, c] for c in range(0, 64, 7))
        recs.append(m0[i])
    recs = np.array(recs)
    assert len(recs) == g["count"] and util.digest(recs) == g["sha256"]


def test_header_present(gpu_ctx2, orc, product_lib):
    import ctypes as C
    import torch
    cfg = B.synth_cfg(600_000, stride=3400, ber=0.03, seed=5, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    pl = util.planted_list(cfg)
    pk = pkts_for(pl, s)
    pk["length"][::5] = 100
    d_s = torch.from_numpy(s).cuda()
    d_p = torch.from_numpy(pk.view(np.uint8)).cuda()
    d_r = torch.zeros(len(pl), dtype=torch.uint8, device="cuda")
    B.check(product_lib.btbb_b200_header_present_dev(gpu_ctx2.h, d_s.data_ptr(), len(s), d_p.data_ptr(),
                                                     len(pl), d_r.data_ptr(), 0))
    torch.cuda.synchronize()
    want = [orc.orc_header_present(s[p.offset:].ctypes.data, int(pk["length"][i])) for i, p in enumerate(pl)]
    assert d_r.cpu().numpy().tolist() == want and 0 < sum(want) < len(want)
