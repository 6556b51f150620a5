"""Hop sequence (groundwork): the oracle's closed-form restatement of gen_hops
(bluetooth_piconet.c:311-362) -- one entry per sequence index, any range on its own, the shape a
one-thread-per-entry kernel needs -- against windows of the reference's 2^27-entry table (where
oracle/_ref exists) and against committed digests of the reference's output.  No product code yet."""
import hashlib
import json
import os

import numpy as np
import pytest

import util

FIXTURE = os.path.join(util.GOLDEN, "hops.json")
AFH = bytes([0xFF, 0x0F, 0xF0, 0xAA, 0x55, 0x01, 0x80, 0xFF, 0xFF, 0x7F])
CASES = {"a96ef25": (0xA96EF25, None), "0123456": (0x0123456, None), "fedcba9_afh": (0xFEDCBA9, AFH)}
WINDOWS = [(0, 1 << 20), ((1 << 26) - (1 << 19), 1 << 20), ((1 << 27) - (1 << 20), 1 << 20), (12345679, 777777)]


def _seq(L, prefix, address, afh, first, n):
    out = np.zeros(n, dtype=np.uint8)
    m = np.frombuffer(afh, dtype=np.uint8).copy() if afh else None
    getattr(L, prefix + "_hop_sequence")(address, m.ctypes.data if afh else None, first, n, out.ctypes.data)
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_hop_sequence_windows(name):
    address, afh = CASES[name]
    O = util.oracle()
    fx = json.load(open(FIXTURE))[name]
    ref_full = _seq(util.ref(), "ref", address, afh, 0, 1 << 27) if util.have_ref() else None
    for first, n in WINDOWS:
        got = _seq(O, "orc", address, afh, first, n)
        assert hashlib.sha256(got.tobytes()).hexdigest() == fx[f"{first}+{n}"]
        if ref_full is not None:
            assert np.array_equal(got, ref_full[first:first + n])
    if ref_full is not None:
        assert hashlib.sha256(ref_full.tobytes()).hexdigest() == fx["full"]


@pytest.mark.skipif(not util.have_ref(), reason="compiled reference not built here")
def test_reference_winnow_agrees_with_a_filter_over_its_table():
    """The harness tests/test_gpu_hops.py leans on: the reference's btbb_init_hop_reversal + btbb_winnow
    (oracle/ref_shim.c:ref_hop_winnow) give the candidate counts a plain filter over the sequence gives."""
    import ctypes as C
    R = util.ref()
    R.ref_hop_winnow.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    address, afh = CASES["a96ef25"]
    full = _seq(util.oracle(), "orc", address, afh, 0, 1 << 27)
    mask = (1 << 27) - 1
    true_clk = 0x2345678
    idx = np.array([0, 7, 9, 30, 31, 60, 100], dtype=np.int32)
    ch = np.ascontiguousarray(full[(true_clk + idx) & mask])
    counts = np.zeros(len(idx), dtype=np.int32)
    rc = np.zeros(1 << 16, dtype=np.uint32)
    n = R.ref_hop_winnow(address, None, 0, true_clk & 63, len(idx), idx.ctypes.data, ch.ctypes.data, counts.ctypes.data, rc.ctypes.data, len(rc))
    want = np.arange(true_clk & 63, 1 << 27, 64, dtype=np.int64)
    traj = []
    for j in range(len(idx)):
        want = want[full[(want + int(idx[j])) & mask] == ch[j]]
        traj.append(len(want))
    stop = next(j for j, v in enumerate(traj) if v <= 1)
    assert counts[:stop + 1].tolist() == traj[:stop + 1] and n == 1 and rc[0] == true_clk
