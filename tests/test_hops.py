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
