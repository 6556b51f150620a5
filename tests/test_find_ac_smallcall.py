"""The host short-search path of the classic btbb_find_ac (find_ac_host.cpp: packed sliding window,
table-driven syndrome, open-addressing error table) through btbb_b200_find_first_smallcall, against
the oracle -- no GPU involved.  tests/test_gpu_compat.py drives the same routine through btbb_find_ac
with the context's own host copy of the table."""
import ctypes as C

import numpy as np
import pytest

import util
from util import B


def _walk(lib, s, n, lap, table_k, k, limit=10**9):
    """every hit of the stream, found the way a classic caller does: search, step past the hit, search again"""
    out, pos = [], 0
    hit, found = B.Hit(), C.c_int(0)
    while pos < n and len(out) < limit:
        B.check(lib.btbb_b200_find_first_smallcall(s.ctypes.data + pos, n - pos, lap, table_k, k, C.byref(hit), C.byref(found)))
        if not found.value:
            break
        out.append((pos + hit.offset, hit.lap, hit.ac_errors))
        pos += hit.offset + 1
    return out


@pytest.mark.parametrize("table_k", [0, 1, 2, 3, 4])
def test_promiscuous_first_hits_equal_the_oracle_list(product_lib, orc, table_k):
    cfg = B.synth_cfg(260_000, stride=2600, ber=0.012, seed=1000 + table_k, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    n = len(s) - 63
    # the oracle, like the reference, builds its error table once per loaded library (for the first
    # non-zero k, bluetooth_packet.c:288-289): a private copy of the library per table size
    import os
    import shutil
    util.oracle()
    tmp = f"/tmp/liboracle_k{table_k}_{os.getpid()}.so"
    shutil.copy(util.ORACLE_SO, tmp)
    mine = C.CDLL(tmp)
    os.remove(tmp)
    mine.orc_find_all.restype = C.c_int64
    mine.orc_find_all.argtypes = orc.orc_find_all.argtypes
    mine.orc_init(table_k)
    for k in sorted({0, table_k, min(table_k + 1, 5)}):
        want = util.find_all(mine, "orc", s, n, B.LAP_ANY, k)
        got = _walk(product_lib, s, n, B.LAP_ANY, table_k, k)
        assert got == [(int(h["offset"]), int(h["lap"]), int(h["ac_errors"])) for h in want], (table_k, k)
        assert len(got) > (20 if k == 0 else 60) or table_k == 0


def test_known_lap_and_edges(product_lib, orc):
    cfg = B.synth_cfg(120_000, stride=2000, ber=0.02, seed=77, n_laps=3, mix=("ID", "DM1", "FHS"))
    s = B.synth_host(cfg)
    n = len(s) - 63
    laps = sorted({p.lap for p in util.planted_list(cfg)})
    for lap in laps:
        for k in (0, 1, 3, 6, 10):
            want = util.find_all(orc, "orc", s, n, lap, k)
            assert _walk(product_lib, s, n, lap, 0, k) == [(int(h["offset"]), int(h["lap"]), int(h["ac_errors"])) for h in want], (lap, k)
    # lengths around the packing granularity and the 4096-position blocks; a hit exactly at the last position
    p = util.planted_list(cfg)[1]
    hit, found = B.Hit(), C.c_int(0)
    for extra in (0, 1, 7, 8, 63, 64, 65, 4095, 4096, 4097):
        start = max(0, p.offset - extra)
        for length in (p.offset - start, p.offset - start + 1):
            B.check(product_lib.btbb_b200_find_first_smallcall(s.ctypes.data + start, length, p.lap, 0, 0, C.byref(hit), C.byref(found)))
            prior = [h for h in util.find_all(orc, "orc", s[start:], max(length, 0), p.lap, 0)] if length > 0 else []
            assert found.value == (1 if prior else 0), (extra, length)
            if prior:
                assert hit.offset == int(prior[0]["offset"])
    B.check(product_lib.btbb_b200_find_first_smallcall(s.ctypes.data, 0, B.LAP_ANY, 2, 2, C.byref(hit), C.byref(found)))
    assert found.value == 0
    assert product_lib.btbb_b200_find_first_smallcall(s.ctypes.data, 10, B.LAP_ANY, 5, 2, C.byref(hit), C.byref(found)) == -1
