"""The synthetic capture generator (host half) against the oracle: every planted packet is
found at its offset and decodes with the expected result.  CPU only."""
import numpy as np

import util
from util import B


def test_planted_packets_round_trip(product_lib, orc):
    cfg = B.synth_cfg(1_000_000, stride=5000, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    assert orc.orc_init(2) == 0
    hits = util.find_all(orc, "orc", s, len(s) - 63, B.LAP_ANY, 0)
    by_off = {int(h["offset"]): h for h in hits}
    expect = {"ID": (0, 0), "DM1": (1, 10), "DH1": (1, 10), "DM3": (1, 10), "FHS": (1, 1000), "HV1": (1, 2),
              "DM5": (1, 10), "DH3": (1, 10)}
    names = {v: k for k, v in B.KIND.items()}
    seen = set()
    for p in util.planted_list(cfg):
        assert p.offset in by_off and by_off[p.offset]["lap"] == p.lap
        d = util.decode_one(orc, "orc", s, p.offset, min(3125, len(s) - p.offset), p.clk6, p.uap)
        assert (d["header_ok"], d["rv"]) == expect[names[p.kind]]
        if p.kind != 0:
            assert d["lt_addr"] == p.lt_addr and d["uap"] == p.uap
        seen.add(p.kind)
    assert seen == set(range(8))


def test_shards_are_slices_of_the_whole(product_lib):
    cfg = B.synth_cfg(300_000, stride=4096, ber=0.01)
    whole = B.synth_host(cfg)
    for first, n in ((0, 1000), (777, 100_001), (123_456, 50_000), (299_000, 1000)):
        part = B.synth_host(B.synth_cfg(n, stride=4096, ber=0.01, first_symbol=first))
        assert np.array_equal(part, whole[first:first + n])


def test_ber_rate(product_lib):
    a = B.synth_host(B.synth_cfg(500_000, stride=0))
    b = B.synth_host(B.synth_cfg(500_000, stride=0, ber=0.02))
    rate = float((a != b).mean())
    assert 0.018 < rate < 0.022
    assert set(np.unique(a).tolist()) == {0, 1}
