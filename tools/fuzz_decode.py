"""Differential fuzz of the per-packet chain on the CPU (no GPU): oracle vs the unmodified reference
(oracle/_ref, where built) and the product's host small-call path vs the oracle, over noise / FEC-clean
packets with arbitrary lengths 0..3125, every single decoder with a forced (matching or foreign) type,
btbb_decode with and without the raw-payload flag, and the 64-clock sweep.

    python tools/fuzz_decode.py [seed] [cases]        # prints the number of mismatches (0 expected)

Run for this round with seeds 1, 2, 3 x 12 000 cases (host path) and 11, 12 x 15 000 cases
(oracle vs reference): 0 mismatches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
from util import B

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
orc = util.oracle()
R = util.ref() if util.have_ref() else None
if R is not None:
    R.btbb_init(2)
rng = np.random.default_rng(seed)
bad = 0


def differ(what, a, b, *ctx):
    global bad
    if a.tobytes() != b.tobytes():
        bad += 1
        if bad < 6:
            print("MISMATCH", what, ctx, a, b)


for i, (sym, n, clk, uap, t, fn, w) in enumerate(util.forced_type_cases(orc, rng, N)):
    if i % 2 == 0:
        n = int(rng.integers(0, 3126))
    for raw in (0, 1):
        want = util.typed_one(orc, "orc", sym, n, clk, uap, t, fn, w, raw)
        if R is not None:
            differ("oracle/reference typed", want, util.typed_one(R, "ref", sym, n, clk, uap, t, fn, w, raw), i, n, clk, uap, t, fn, w, raw)
        got = B.decode_smallcall(sym, n, clk, uap, whitened=w, ptype=t, mode=util.mode_of_fn(fn) | (B.MODE_FLAG_RAW_PAYLOAD if raw else 0))[0]
        differ("host/oracle typed", got, want, i, n, clk, uap, t, fn, w, raw)
    if i % 3 == 0:
        want = util.decode_one_raw(orc, "orc", sym, 0, n, clk, uap, whitened=w)
        if R is not None:
            differ("oracle/reference decode", want, util.decode_one_raw(R, "ref", sym, 0, n, clk, uap, whitened=w), i, n, clk, uap, w)
        differ("host/oracle decode raw", B.decode_smallcall(sym, n, clk, uap, whitened=w, mode=B.MODE_FLAG_RAW_PAYLOAD)[0], want, i, n, clk, uap, w)
        differ("host/oracle decode", B.decode_smallcall(sym, n, clk, uap, whitened=w)[0],
               util.decode_one(orc, "orc", sym, 0, n, clk, uap, whitened=w), i, n, clk, uap, w)
    if i % 16 == 0:
        tc = B.decode_smallcall(sym, n, whitened=w, mode=B.MODE_TRY_CLOCKS)
        for c in range(0, 64, 5):
            want = util.try_clock_one(orc, "orc", sym, 0, n, c, whitened=w)
            if R is not None:
                differ("oracle/reference try_clock", want, util.try_clock_one(R, "ref", sym, 0, n, c, whitened=w), i, n, c, w)
            differ("host/oracle try_clock", tc[c], want, i, n, c, w)
print("cases", N, "seed", seed, "reference", R is not None, "mismatches", bad)
