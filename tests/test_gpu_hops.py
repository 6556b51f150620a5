"""Hop sequence and CLK1-27 winnowing on the GPU: btbb_b200_hop_sequence_dev
against the oracle's closed form and the digests of the reference's own 2^27-entry table
(tests/golden/hops.json), btbb_b200_hop_winnow against a filter over that table and -- where the
compiled reference travels with the tree -- against the reference's btbb_init_hop_reversal /
btbb_winnow (bluetooth_piconet.c:475-499, :613-645) itself."""
import ctypes as C
import hashlib
import json

import numpy as np
import pytest

import util
from util import B
import test_hops

pytestmark = pytest.mark.gpu


def _gpu_seq(ctx, lib, address, afh, first, n, shift=0):
    import torch
    d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    cfg = B.hop_cfg(address, afh)
    B.check(lib.btbb_b200_hop_sequence_dev(ctx.h, C.byref(cfg), first, n, d.data_ptr() + shift, 0))
    torch.cuda.synchronize()
    return d[shift:shift + n].cpu().numpy()


@pytest.mark.parametrize("name", sorted(test_hops.CASES))
def test_hop_sequence_matches_reference_table(gpu_ctx2, product_lib, name):
    address, afh = test_hops.CASES[name]
    fx = json.load(open(test_hops.FIXTURE))[name]
    full = _gpu_seq(gpu_ctx2, product_lib, address, afh, 0, 1 << 27)
    assert hashlib.sha256(full.tobytes()).hexdigest() == fx["full"]          # the reference's whole table
    O = util.oracle()
    for first, n in test_hops.WINDOWS + [(5, 100), (1 << 20, 16), ((1 << 27) - 33, 33)]:
        for shift in (0, 3):      # ranges / output pointers that are not 16-byte aligned
            got = _gpu_seq(gpu_ctx2, product_lib, address, afh, first, n, shift)
            assert np.array_equal(got, full[first:first + n])
        assert np.array_equal(test_hops._seq(O, "orc", address, afh, first, min(n, 4096)), full[first:first + min(n, 4096)])


@pytest.mark.parametrize("name,aliased", [("a96ef25", False), ("fedcba9_afh", False), ("0123456", True)])
def test_hop_winnow(gpu_ctx2, product_lib, name, aliased):
    address, afh = test_hops.CASES[name]
    full = _gpu_seq(gpu_ctx2, product_lib, address, afh, 0, 1 << 27)      # validated against the reference above
    seen = ((full.astype(np.int32) + 24) % 25 + 26).astype(np.uint8) if aliased else full
    rng = np.random.default_rng(5)
    cfg = B.hop_cfg(address, afh, aliased)
    mask = (1 << 27) - 1
    R = util.ref() if util.have_ref() else None
    if R is not None:
        R.ref_hop_winnow.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    for true_clk in (int(rng.integers(0, 1 << 27)), (1 << 27) - 5, 77):
        idx = np.concatenate([[0], np.cumsum(rng.integers(1, 40, 11))]).astype(np.int32)
        ch = seen[(true_clk + idx) & mask]
        known6 = true_clk & 63
        for n_obs in (1, 2, 3, len(idx)):
            cands, after = B.hop_winnow(gpu_ctx2, cfg, known6, idx[:n_obs], ch[:n_obs])
            want = np.arange(known6, 1 << 27, 64, dtype=np.int64)
            traj = []
            for j in range(n_obs):
                want = want[seen[(want + int(idx[j])) & mask] == ch[j]]
                traj.append(len(want))
            assert after.tolist() == traj and np.array_equal(cands.astype(np.int64), want)
            assert true_clk in cands
        assert 1 <= traj[-1] <= 4 and true_clk in cands      # twelve hops (all but) pin the clock down
        if R is not None:
            m = np.frombuffer(afh, dtype=np.uint8).copy() if afh else None
            counts = np.zeros(len(idx), dtype=np.int32)
            rc = np.zeros(1 << 16, dtype=np.uint32)
            n = R.ref_hop_winnow(address, m.ctypes.data if afh else None, int(aliased), known6, len(idx), idx.ctypes.data,
                                 np.ascontiguousarray(ch).ctypes.data, counts.ctypes.data, rc.ctypes.data, len(rc))
            stop = next((j for j, v in enumerate(traj) if v <= 1), len(traj) - 1)      # btbb_winnow breaks at <= 1 (:622)
            assert counts[:stop + 1].tolist() == traj[:stop + 1] and n == traj[stop]
            assert rc[:n].tolist() == cands.tolist() if stop == len(traj) - 1 else rc[0] == true_clk
    # an observation no candidate agrees with: nothing survives
    cands, after = B.hop_winnow(gpu_ctx2, cfg, 5, [0, 1, 2, 3, 4, 5, 6, 7], [3, 3, 3, 3, 3, 3, 3, 3])
    assert len(cands) == 0 and after[-1] == 0
